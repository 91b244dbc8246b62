mkdir -p gpurun_out/cfg
for c in 2 3 4 5; do timeout 600 python scripts/config_bench.py $c 2>gpurun_out/cfg/c${c}_n1.err | tee gpurun_out/cfg/c${c}_n1.json | cut -c1-330; done
tail -3 gpurun_out/cfg/c5_n1.err
