timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for mb in 0 3; do echo MINB=$mb; SPLATT_B200_MINB=$mb timeout 300 python scripts/quick_bench.py 5000 50000000 16 4 0 2>&1 | grep -E 'mode 0'; done
SPLATT_B200_MINB=3 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "t4 or t5 or t6 or t8" 2>&1 | tail -2
