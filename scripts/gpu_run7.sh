mkdir -p gpurun_out/r2
echo "=== multi tests on 2 GPUs"
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_config1_cli.py -x -q -m gpu > gpurun_out/r2/pytest_multi_2gpu.log 2>&1; echo "pytest multi rc=$?"
tail -8 gpurun_out/r2/pytest_multi_2gpu.log
echo "=== drift check"
timeout 300 python scripts/debug_multi_cpd.py 300 200000 32 16 0,0 0,1 2>&1 | grep -v "its ="
echo "=== test_fused N=2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/test_fused.py > gpurun_out/r2/fused2.log 2>&1; echo "fused rc=$?"
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2/fused2.log | tail -6
echo "=== bench N=2"
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err; echo "bench rc=$?"
tail -4 gpurun_out/r2/bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_n2.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","parity_rel_fro","step_ms_min","step_ms_max"):
        print(k, d.get(k))
    e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","error")})
    for k,v in (d.get("named_configs") or {}).items():
        print("named",k, {kk:v.get(kk) for kk in ("ms_per_step","per_mode_ms","parity_rel_fro","error")})
    print("cpd", d["cpd_als_iteration"])
except Exception as e:
    print("parse failed", e)
    print(open("gpurun_out/r2/bench_n2.json").read()[-3000:])
PY
