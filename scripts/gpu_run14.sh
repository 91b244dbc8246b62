mkdir -p gpurun_out/r2
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --named "") > gpurun_out/r2/bench_n2.out 2> gpurun_out/r2/bench_n2.err; echo "bench rc=$?"
tail -1 gpurun_out/r2/bench_n2.out > gpurun_out/r2/bench_n2.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_n2.json").read())
for k in ("value","ms_per_step","parity_rel_fro"): print(k, d.get(k))
e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","error")})
print("cpd", {k:v for k,v in d["cpd_als_iteration"].items() if k in ("ours_ms","c_abi_ms","c_abi_fit")})
PY
