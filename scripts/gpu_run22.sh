mkdir -p gpurun_out/r2
(time NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5) > gpurun_out/r2/bench_n8.out 2> gpurun_out/r2/bench_n8.err; echo "bench rc=$?"
tail -2 gpurun_out/r2/bench_n8.err
grep -c "Init COMPLETE" gpurun_out/r2/bench_n8.out
tail -1 gpurun_out/r2/bench_n8.out > gpurun_out/r2/bench_n8.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_n8.json").read())
    for k in ("value","ms_per_step","parity_rel_fro","step_ms_min","step_ms_max","clocks"): print(k, d.get(k))
    e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","error")})
    for k,v in (d.get("named_configs") or {}).items():
        print("named",k, {kk:v.get(kk) for kk in ("ms_per_step","per_mode_ms","parity_rel_fro","error")})
    print("cpd", {k:v for k,v in d["cpd_als_iteration"].items() if k in ("ours_ms","c_abi_ms","c_abi_fit")})
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2/bench_n8.out").read()[-1500:])
PY
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2/pytest_multi_8gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2/pytest_multi_8gpu.log
