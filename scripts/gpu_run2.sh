set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/r2/pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -15 gpurun_out/r2/pytest_multi.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_multi_gpu.py > gpurun_out/r2/pytest2.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2/pytest2.log
timeout 300 python scripts/probe_sweep.py 10000 32 20000000 > gpurun_out/r2/probe_R32.json 2> gpurun_out/r2/probe_R32.err
timeout 300 python scripts/probe_sweep.py 5000 16 60000000 > gpurun_out/r2/probe_R16.json 2> gpurun_out/r2/probe_R16.err
python - <<'PY'
import json
for f in ("gpurun_out/r2/probe_R32.json","gpurun_out/r2/probe_R16.json"):
    try:
        d=json.load(open(f)); print(f, d["best"])
        for r in sorted(d["sweep"], key=lambda r:-r["TBps"])[:6]: print("   ", r)
        for r in d["l1_squeeze_3ctas_8rows"]: print("  L1", r["smem_per_cta"], round(r["TBps"],2))
    except Exception as e: print(f, "failed", e)
PY
# ncu: 4-mode root kernel as shipped (config 3), 3-mode root (config 2)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mttkrp_stream -s 8 -c 1 -o gpurun_out/r2/prof_root4_anc python scripts/quick_bench.py 5000 50000000 16 4 > gpurun_out/r2/ncu_root4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mttkrp_stream -s 6 -c 1 -o gpurun_out/r2/prof_root3 python scripts/quick_bench.py 10000 10000000 32 3 > gpurun_out/r2/ncu_root3.log 2>&1
ls -la gpurun_out/r2
