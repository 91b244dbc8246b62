CSF_ALLOC=0 timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 1 2>&1 | grep -E 'mode'
CSF_ALLOC=0 ncu --set full --clock-control none -k regex:mttkrp_stream -s 9 -c 6 -f -o gpurun_out/prof_intl_leaf python scripts/quick_bench.py 10000 10000000 32 3 1 > gpurun_out/ncu_il.log 2>&1; tail -1 gpurun_out/ncu_il.log
