mkdir -p gpurun_out/r2
nvidia-smi -L
echo "=== multi tests on 2 GPUs"
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/r2/pytest_multi_2gpu.log 2>&1; echo "pytest multi rc=$?"
tail -25 gpurun_out/r2/pytest_multi_2gpu.log
echo "=== new single-gpu tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pinned_shared or pageable_paths or reuses_device_mirror or rank_128" > gpurun_out/r2/pytest_new.log 2>&1; echo "pytest new rc=$?"
tail -15 gpurun_out/r2/pytest_new.log
echo "=== test_fused N=2 (in-kernel barrier)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/test_fused.py > gpurun_out/r2/fused2.log 2>&1; echo "fused rc=$?"
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2/fused2.log | tail -15
echo "=== bench N=2"
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err; echo "bench rc=$?"
tail -12 gpurun_out/r2/bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_n2.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","parity_rel_fro","step_ms_min","step_ms_max","gpu_launches","clocks"):
        print(k, d.get(k))
    print("exchange", d["config"]["exchange"])
    print("e2e", d["e2e"])
    print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","launch_ms","per_mode_ms")})
    for k,v in (d.get("named_configs") or {}).items():
        print("named",k, {kk:v.get(kk) for kk in ("ms_per_step","per_mode_ms","value","parity_rel_fro","clocks","exchange","error")})
    print("cpd", d["cpd_als_iteration"])
except Exception as e:
    print("parse failed", e)
    print(open("gpurun_out/r2/bench_n2.json").read()[-3000:])
PY
