mkdir -p gpurun_out/r2
echo "=== full gpu tests"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_full.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2/pytest_full.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench N=1 (driver command line)"
(time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2/bench_n1.out 2> gpurun_out/r2/bench_n1.err; echo "bench rc=$?"
tail -1 gpurun_out/r2/bench_n1.out > gpurun_out/r2/bench_n1.json
tail -3 gpurun_out/r2/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_n1.json").read())
for k in ("value","ms_per_step","warmup","steps","parity_rel_fro","clocks","gpu_launches","exchange"): print(k, d.get(k))
e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","h2d_bytes_per_step","d2h_bytes_per_step","error")})
print("roofline", {k:d["roofline"][k] for k in ("achieved","peak","frac","launch_ms","traffic","per_mode_ms")})
print("gather", d["roofline"]["gather_path"])
print("cpu", d["cpu_baseline"])
for k,v in (d.get("named_configs") or {}).items():
    print("named",k, {kk:v.get(kk) for kk in ("ms_per_step","per_mode_ms","parity_rel_fro","clocks","error")})
print("cpd", d["cpd_als_iteration"])
PY
echo "=== reference arm"
(time timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2) > gpurun_out/r2/bench_ref.json 2> gpurun_out/r2/bench_ref.err; echo "ref rc=$?"
tail -3 gpurun_out/r2/bench_ref.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/r2/bench_ref.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','impl')}, d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:120])"
