timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/test_fused.py 2>&1 | grep -v -E '^\*|OMP_NUM|^$' | tail -25
