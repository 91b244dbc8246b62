mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_multi_process.py tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2/pytest_multi_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/pytest_multi_2gpu.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/r2/bench_n2.out 2> gpurun_out/r2/bench_n2.err; echo "bench rc=$?"
tail -1 gpurun_out/r2/bench_n2.out > gpurun_out/r2/bench_n2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_n2.json").read())
for k in ("value","ms_per_step","parity_rel_fro","clocks"): print(k, d.get(k))
e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","error")})
for k,v in (d.get("named_configs") or {}).items():
    print("named",k, {kk:v.get(kk) for kk in ("per_mode_ms","parity_rel_fro","error")})
print("cpd", {k:v for k,v in d["cpd_als_iteration"].items() if k in ("ours_ms","c_abi_ms","c_abi_fit")})
PY
