timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/test_fused.py 2>&1 | grep -v -E '^\*|OMP_NUM|^$' | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 50 --warmup 3 2> gpurun_out/bench_n8.err | tee gpurun_out/bench_n8.json | cut -c1-400
tail -2 gpurun_out/bench_n8.err
