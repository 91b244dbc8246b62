mkdir -p gpurun_out/r2
echo "=== cpd tests"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "cpd" > gpurun_out/r2/pytest_cpd.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2/pytest_cpd.log
K='regex:k_gram|k_form_chol|k_solve_rows|k_colnorm|k_finish_lambda|k_scale_cols|k_inner|mttkrp_stream'
for cfg in "10000 10000 10000 10000000 32 cfg2" "1000000 1000000 1000 50000000 64 cfg5"; do
  set -- $cfg
  echo "=== cpd launch list $6"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2/cpd_launches_$6_v2.csv python scripts/cpd_profile.py $1 $2 $3 $4 $5 3 > gpurun_out/r2/cpd_$6_v2.log 2>&1
  tail -1 gpurun_out/r2/cpd_$6_v2.log
done
python - <<'PY'
import csv, collections, re
for f in ("gpurun_out/r2/cpd_launches_cfg2_v2.csv","gpurun_out/r2/cpd_launches_cfg5_v2.csv"):
    try:
        rows=[r for r in csv.reader(open(f)) if len(r)>5]
        hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
        tot=collections.defaultdict(float); cnt=collections.Counter()
        for r in rows[1:]:
            v=float(r[vi].replace(",","")); u=r[ui]
            if u=="ns": v/=1e3
            elif u=="ms": v*=1e3
            m=re.search(r"(k_\w+|mttkrp_stream_kernel)", r[ki]); name=m.group(1) if m else r[ki][:30]
            tot[name]+=v; cnt[name]+=1
        T=sum(tot.values()); print(f, "total us", round(T))
        for k,v in sorted(tot.items(), key=lambda kv:-kv[1]): print(f"   {k:24s} {cnt[k]:4d} launches {v:10.1f} us  {100*v/T:5.1f} %   {v/cnt[k]:9.1f} us each")
    except Exception as e: print(f, "failed", e)
PY
