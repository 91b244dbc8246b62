mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_multi_process.py tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2/pytest_multi_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/pytest_multi_2gpu.log
for st in 1 0; do
echo "=== MC_STORE=$st"
SPLATT_B200_MC_STORE=$st timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$st bench.py --gpus 2 --steps 20 --warmup 5 --named 5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('value','ms_per_step','parity_rel_fro'): print(k, d.get(k))
v=d['named_configs']['5']; print('named 5', v['per_mode_ms'], v['parity_rel_fro'])
e=d['e2e']; print('e2e', e['ms_per_step'], e['pinned']['ms_per_step'], e.get('parity_rel_fro_mode0_vs_device_path'))
print('cpd', {k:v for k,v in d['cpd_als_iteration'].items() if k in ('ours_ms','c_abi_ms','c_abi_fit')})"
done
