mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_config1_cli.py -x -q -m gpu > gpurun_out/r2/pytest_multi_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/pytest_multi_2gpu.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --named "") > gpurun_out/r2/bench_n2.out 2> gpurun_out/r2/bench_n2.err; echo "bench rc=$?"
tail -1 gpurun_out/r2/bench_n2.out > gpurun_out/r2/bench_n2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_n2.json").read())
for k in ("value","ms_per_step","parity_rel_fro"): print(k, d.get(k))
e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","parity_rel_fro_mode0_vs_device_path","error")})
PY
echo "--- pipeline off"
SPLATT_B200_PIPELINE=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --named "" 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('e2e', {k:e.get(k) for k in ('ms_per_step','pinned','pageable_over_pinned')})"
