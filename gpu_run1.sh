nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; lscpu | grep 'Model name'
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -30
timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 0 2>&1 | tail -12
timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 1 2>&1 | tail -12
timeout 300 python scripts/quick_bench.py 5000 50000000 16 4 0 2>&1 | tail -12
