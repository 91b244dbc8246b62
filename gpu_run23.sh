timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cta_tiled" 2>&1 | tail -5
KTILE=0 ncu --set full --clock-control none --import-source on -k regex:mttkrp_tiled -s 3 -c 1 -f -o gpurun_out/prof_tiled python scripts/quick_bench.py 10000 10000000 32 3 0 > gpurun_out/ncu_tiled.log 2>&1; tail -1 gpurun_out/ncu_tiled.log
