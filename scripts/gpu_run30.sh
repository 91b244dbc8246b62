timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 --named 5 2>/dev/null | tail -1 > gpurun_out/bench_n4_mcstore.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n4_mcstore.json").read())
for k in ('value','ms_per_step','parity_rel_fro','step_ms_min','step_ms_max'): print(k, d.get(k))
v=d['named_configs']['5']; print('named 5', v['per_mode_ms'], v['parity_rel_fro'], v['clocks'])
e=d['e2e']; print('e2e', e['ms_per_step'], e['pinned']['ms_per_step'])
print('cpd', {k:v for k,v in d['cpd_als_iteration'].items() if k in ('ours_ms','c_abi_ms','c_abi_fit')})
PY
