timeout 300 python scripts/quick_bench.py 5000 50000000 16 4 0 2>&1 | grep mode
timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 0 2>&1 | grep mode
timeout 300 python scripts/quick_bench.py 1000000 25000000 64 3 0 zipf 2>&1 | grep -E 'mode|stream'
timeout 300 python scripts/quick_bench.py 1000000 25000000 64 3 1 zipf 2>&1 | grep -E 'mode|stream'
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for t in 32 64 128; do SPLATT_REF_THREADS=$t python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ref threads', d['cpu_baseline']['cores'], d['value']/1e9, 'G')"; done
