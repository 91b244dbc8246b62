timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cta_tiled" 2>&1 | tail -8
for kt in -1 0; do echo KTILE=$kt; KTILE=$kt timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 0 2>&1 | grep -E 'mode|stream 0'; done
KTILE=0 timeout 300 python scripts/quick_bench.py 10000 10000000 16 3 0 2>&1 | grep -E 'mode 0|stream 0'
KTILE=0 timeout 300 python scripts/quick_bench.py 10000 10000000 64 3 0 2>&1 | grep -E 'mode 0|stream 0'
