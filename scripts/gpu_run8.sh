mkdir -p gpurun_out/r2
echo "=== full gpu tests"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_full.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2/pytest_full.log
K='regex:k_gram|k_form_chol|k_solve_rows|k_colnorm|k_finish_lambda|k_scale_cols|k_inner|mttkrp_stream'
echo "=== cpd launch list config 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2/cpd_launches_cfg2.csv python scripts/cpd_profile.py 10000 10000 10000 10000000 32 3 > gpurun_out/r2/cpd_cfg2.log 2>&1
tail -2 gpurun_out/r2/cpd_cfg2.log
echo "=== cpd launch list config-5 shape (1M x 1M x 1K, 50M nnz, R=64)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2/cpd_launches_cfg5.csv python scripts/cpd_profile.py 1000000 1000000 1000 50000000 64 3 > gpurun_out/r2/cpd_cfg5.log 2>&1
tail -2 gpurun_out/r2/cpd_cfg5.log
python - <<'PY'
import csv, collections
for f in ("gpurun_out/r2/cpd_launches_cfg2.csv","gpurun_out/r2/cpd_launches_cfg5.csv"):
    try:
        rows=[r for r in csv.reader(open(f)) if len(r)>5]
        hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
        tot=collections.defaultdict(float); cnt=collections.Counter()
        for r in rows[1:]:
            v=float(r[vi].replace(",","")); u=r[ui]
            if u=="ns": v/=1e3
            elif u=="ms": v*=1e3
            elif u in ("s","second"): v*=1e6
            name=r[ki].split("(")[0].split("<")[0]
            tot[name]+=v; cnt[name]+=1
        T=sum(tot.values()); print(f, "total us", round(T))
        for k,v in sorted(tot.items(), key=lambda kv:-kv[1]): print(f"   {k:28s} {cnt[k]:4d} launches {v:10.1f} us  {100*v/T:5.1f} %")
    except Exception as e: print(f, "failed", e)
PY
echo "=== bench N=1"
(time timeout 1200 python bench.py --steps 20 --warmup 5) > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_n1.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","parity_rel_fro"): print(k, d.get(k))
e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","error")})
print("cpd", d["cpd_als_iteration"])
PY
