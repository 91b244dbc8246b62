mkdir -p gpurun_out/r2
echo "=== multi tests + CLI on 2 GPUs"
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_config1_cli.py -x -q -m gpu > gpurun_out/r2/pytest_multi_2gpu.log 2>&1; echo "pytest multi rc=$?"
tail -4 gpurun_out/r2/pytest_multi_2gpu.log
echo "=== config2 variants"
for v in "0 0" "3 4" "3 3"; do
  set -- $v
  echo "--- BATCH=$1 MINB=$2"
  SPLATT_B200_BATCH=$1 SPLATT_B200_MINB=$2 timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 2>&1 | grep "^mode [01]"
done
echo "=== long-fiber context: 10K x 10K x 64, 10M nnz, R=32"
DIMS=10000,10000,64 timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 2>&1 | grep "^mode"
echo "=== long-fiber context: 1000 x 1000 x 10000, 10M nnz, R=32"
DIMS=1000,1000,10000 timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 2>&1 | grep "^mode"
echo "=== bench N=2"
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/r2/bench_n2.out 2> gpurun_out/r2/bench_n2.err; echo "bench rc=$?"
tail -2 gpurun_out/r2/bench_n2.err
tail -1 gpurun_out/r2/bench_n2.out > gpurun_out/r2/bench_n2.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_n2.json").read())
    for k in ("value","ms_per_step","parity_rel_fro","exchange"):
        print(k, d.get(k))
    e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","error")})
    print("cpd", d["cpd_als_iteration"])
except Exception as e:
    print("parse failed", e); print(open("gpurun_out/r2/bench_n2.out").read()[-2000:])
PY
