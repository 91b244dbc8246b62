set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/r2/pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -15 gpurun_out/r2/pytest_multi.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_multi_gpu.py > gpurun_out/r2/pytest3.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest3.log
for v in "0 0" "2 4" "4 4"; do
  set -- $v
  echo "=== config2 BATCH=$1 MINB=$2"
  SPLATT_B200_BATCH=$1 SPLATT_B200_MINB=$2 timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 2>&1 | grep -v "^SPLATT-B200"
done > gpurun_out/r2/cfg2_variants.log 2>&1
grep -v "^+" gpurun_out/r2/cfg2_variants.log
for v in "0 0" "3 3" "2 4" "4 0"; do
  set -- $v
  echo "=== config3 BATCH=$1 MINB=$2"
  SPLATT_B200_BATCH=$1 SPLATT_B200_MINB=$2 timeout 300 python scripts/quick_bench.py 5000 50000000 16 4 2>&1 | grep -v "^SPLATT-B200"
done > gpurun_out/r2/cfg3_variants2.log 2>&1
grep -v "^+" gpurun_out/r2/cfg3_variants2.log
