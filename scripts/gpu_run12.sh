mkdir -p gpurun_out/r2
nvidia-smi -L | wc -l
echo "=== bench N=8 (NCCL_DEBUG=INFO as the driver sets it)"
(time NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5) > gpurun_out/r2/bench_n8.out 2> gpurun_out/r2/bench_n8.err; echo "bench rc=$?"
tail -3 gpurun_out/r2/bench_n8.err
grep -c "NCCL INFO" gpurun_out/r2/bench_n8.out; grep "Init COMPLETE" gpurun_out/r2/bench_n8.out | head -3 | cut -c1-200
tail -1 gpurun_out/r2/bench_n8.out > gpurun_out/r2/bench_n8.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_n8.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","parity_rel_fro","step_ms_min","step_ms_max","clocks"):
        print(k, d.get(k))
    print("exchange", d["config"]["exchange"])
    e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","error")})
    print("roofline per_mode", d["roofline"]["per_mode_ms"])
    for k,v in (d.get("named_configs") or {}).items():
        print("named",k, {kk:v.get(kk) for kk in ("ms_per_step","per_mode_ms","value","parity_rel_fro","clocks","error")})
    print("cpd", d["cpd_als_iteration"])
except Exception as e:
    print("parse failed", e)
    print(open("gpurun_out/r2/bench_n8.out").read()[-3000:])
PY
echo "=== multi tests on 8 GPUs"
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/r2/pytest_multi_8gpu.log 2>&1; echo "pytest multi rc=$?"
tail -5 gpurun_out/r2/pytest_multi_8gpu.log
echo "=== config-5 shape CPD on 1 and 8 GPUs"
timeout 900 python scripts/cpd_config5_shape.py 50000000 1 8 2>&1 | tail -4
