mkdir -p gpurun_out/cfg
for c in 4 5; do
 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2955$c scripts/config_bench.py $c 2>gpurun_out/cfg/c${c}_n8.err | grep '^{' | tee gpurun_out/cfg/c${c}_n8_fused.json | cut -c1-300
 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2956$c scripts/config_bench.py $c --nccl 2>>gpurun_out/cfg/c${c}_n8.err | grep '^{' | tee gpurun_out/cfg/c${c}_n8_nccl.json | cut -c1-300
done
tail -3 gpurun_out/cfg/c5_n8.err
