echo "### MC async"; timeout 300 python scripts/debug_multi_cpd.py 300 200000 32 6 0,1 2>&1 | tail -9
echo "### MC sync after kernels (1)"; SPLATT_B200_MULTI_SYNC=1 timeout 300 python scripts/debug_multi_cpd.py 300 200000 32 6 0,1 2>&1 | tail -2
echo "### MC sync after tails (2)"; SPLATT_B200_MULTI_SYNC=2 timeout 300 python scripts/debug_multi_cpd.py 300 200000 32 6 0,1 2>&1 | tail -2
echo "### MC sync both (3)"; SPLATT_B200_MULTI_SYNC=3 timeout 300 python scripts/debug_multi_cpd.py 300 200000 32 6 0,1 2>&1 | tail -2
echo "### peer-reduce fallback on 2 real GPUs"; SPLATT_B200_MULTICAST=0 timeout 300 python scripts/debug_multi_cpd.py 300 200000 32 6 0,1 2>&1 | tail -2
echo "### MC async R=16"; timeout 300 python scripts/debug_multi_cpd.py 300 200000 16 6 0,1 2>&1 | tail -2
echo "### MC async R=64"; timeout 300 python scripts/debug_multi_cpd.py 300 200000 64 6 0,1 2>&1 | tail -2
