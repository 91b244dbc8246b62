timeout 2400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python scripts/cpd_bench.py 2>&1 | grep -E 'its|ours|reference'
