ncu --set full --clock-control none --import-source on -k regex:mttkrp_stream -s 3 -c 1 -f -o gpurun_out/prof_root_v2 python scripts/quick_bench.py 10000 10000000 32 3 0 > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
