ncu --set full --clock-control none --import-source on -k regex:mttkrp_stream -s 3 -c 1 -f -o gpurun_out/prof_root4 python scripts/quick_bench.py 5000 50000000 16 4 0 > gpurun_out/ncu_full4.log 2>&1
tail -1 gpurun_out/ncu_full4.log
KTILE=256 ncu --set full --clock-control none -k regex:mttkrp_stream -s 3 -c 1 -f -o gpurun_out/prof_root3_kt python scripts/quick_bench.py 10000 10000000 32 3 0 > gpurun_out/ncu_full_kt.log 2>&1
tail -1 gpurun_out/ncu_full_kt.log
