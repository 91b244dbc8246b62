timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 0 2>&1 | grep mode
SPLATT_B200_BATCH=8 timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 0 2>&1 | grep mode
timeout 300 python scripts/quick_bench.py 5000 50000000 16 4 0 2>&1 | grep mode
SPLATT_B200_BATCH=8 timeout 300 python scripts/quick_bench.py 5000 50000000 16 4 0 2>&1 | grep mode
timeout 300 python scripts/quick_bench.py 10000 10000000 64 3 0 2>&1 | grep mode
timeout 300 python scripts/quick_bench.py 10000 10000000 16 3 0 2>&1 | grep mode
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
