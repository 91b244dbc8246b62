mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -x -q -k "cpd" > gpurun_out/r2/pytest_cpd.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2/pytest_cpd.log
K='regex:k_gram|k_form_chol|k_solve_rows|k_pack_chol|k_colnorm|k_finish_lambda|k_scale_cols|k_inner|mttkrp_stream'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/r2/cpd_launches_cfg5_v3.csv python scripts/cpd_profile.py 1000000 1000000 1000 50000000 64 3 > gpurun_out/r2/cpd_cfg5_v3.log 2>&1
tail -1 gpurun_out/r2/cpd_cfg5_v3.log
python - <<'PY'
import csv, collections, re
f="gpurun_out/r2/cpd_launches_cfg5_v3.csv"
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
PY
