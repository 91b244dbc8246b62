set -x
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest1.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest1.log
for v in "0 0" "2 0" "3 3" "3 2" "4 0" "4 3"; do
  set -- $v
  echo "=== config3 BATCH=$1 MINB=$2"
  SPLATT_B200_BATCH=$1 SPLATT_B200_MINB=$2 timeout 300 python scripts/quick_bench.py 5000 50000000 16 4 2>&1 | grep -v "^SPLATT-B200"
done > gpurun_out/r2/cfg3_variants.log 2>&1
cat gpurun_out/r2/cfg3_variants.log
timeout 300 python scripts/probe_sweep.py 10000 32 20000000 > gpurun_out/r2/probe_R32.json 2> gpurun_out/r2/probe_R32.err
timeout 300 python scripts/probe_sweep.py 5000 16 60000000 > gpurun_out/r2/probe_R16.json 2> gpurun_out/r2/probe_R16.err
python - <<'PY'
import json
for f in ("gpurun_out/r2/probe_R32.json","gpurun_out/r2/probe_R16.json"):
    try:
        d=json.load(open(f)); print(f, d["best"])
        for r in sorted(d["sweep"], key=lambda r:-r["TBps"])[:8]: print("   ", r)
    except Exception as e: print(f, "failed", e)
PY
