mkdir -p gpurun_out/r2
for st in 0 1 2; do for v in "0 0" "2 4" "4 0"; do
  set -- $v
  echo "=== config3 STAGGER=$st BATCH=$1 MINB=$2"
  SPLATT_B200_STAGGER=$st SPLATT_B200_BATCH=$1 SPLATT_B200_MINB=$2 timeout 300 python scripts/quick_bench.py 5000 50000000 16 4 2>&1 | grep "^mode [01]"
done; done > gpurun_out/r2/cfg3_stagger.log 2>&1
cat gpurun_out/r2/cfg3_stagger.log
for st in 0 1; do for v in "0 0" "2 4"; do
  set -- $v
  echo "=== config2 STAGGER=$st BATCH=$1 MINB=$2"
  SPLATT_B200_STAGGER=$st SPLATT_B200_BATCH=$1 SPLATT_B200_MINB=$2 timeout 300 python scripts/quick_bench.py 10000 10000000 32 3 2>&1 | grep "^mode [01]"
done; done > gpurun_out/r2/cfg2_stagger.log 2>&1
cat gpurun_out/r2/cfg2_stagger.log
M=l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__registers_per_thread
for st in 0 1 2; do
  echo "=== ncu config3 STAGGER=$st"
  SPLATT_B200_STAGGER=$st timeout 300 ncu --metrics $M --clock-control none -k regex:mttkrp_stream -s 8 -c 1 python scripts/quick_bench.py 5000 50000000 16 4 2>&1 | grep -E "l1tex|lts__|gpu__time|issue_active|occupancy|registers"
done > gpurun_out/r2/ncu_stagger.log 2>&1
cat gpurun_out/r2/ncu_stagger.log
echo "=== bench N=1"
(time timeout 1200 python bench.py --steps 20 --warmup 5) > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2/bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_n1.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","parity_rel_fro","step_ms_min","step_ms_max","gpu_launches","clocks"):
        print(k, d.get(k))
    print("e2e", d["e2e"])
    print("roofline", {k:d["roofline"][k] for k in ("achieved","frac","launch_ms","per_mode_ms")})
    print("cpu", d["cpu_baseline"])
    for k,v in (d.get("named_configs") or {}).items():
        print("named",k, {kk:v.get(kk) for kk in ("ms_per_step","per_mode_ms","value","parity_rel_fro","clocks","error")})
    print("cpd", d["cpd_als_iteration"])
except Exception as e:
    print("parse failed", e)
PY
