mkdir -p gpurun_out/r2
echo "=== multi tests on 2 GPUs"
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_config1_cli.py -x -q -m gpu > gpurun_out/r2/pytest_multi_2gpu.log 2>&1; echo "pytest multi rc=$?"
tail -8 gpurun_out/r2/pytest_multi_2gpu.log
echo "=== drift check (partitioned tail on 0,1; device-0 tail on 0,0)"
timeout 300 python scripts/debug_multi_cpd.py 300 200000 32 16 0,0 0,1 2>&1 | grep -v "its ="
echo "=== drift check, partitioned tail off"
SPLATT_B200_PARTITIONED_TAIL=0 timeout 300 python scripts/debug_multi_cpd.py 300 200000 32 16 0,1 2>&1 | grep -v "its ="
echo "=== config-5 shape CPD iteration: 1 GPU vs 2 GPUs (partitioned / device-0 tail)"
cat > /tmp/cpd5.py <<'PY'
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import splatt_b200 as S
from splatt_b200 import _abi as A
dims=[1000000,1000000,1000]; nnz=50_000_000; R=64
g=torch.Generator(device="cuda").manual_seed(4)
ind=[torch.randint(0,d,(nnz,),device="cuda",dtype=torch.int32,generator=g) for d in dims]
for m,d in enumerate(dims):
    n=min(d,nnz); ind[m][:n]=torch.arange(n,device="cuda",dtype=torch.int32)
vals=torch.rand(nnz,device="cuda",dtype=torch.float64,generator=g)
o=S.default_opts()
csf=S.csf_alloc(dims,[i.cpu().numpy() for i in ind],vals.cpu().numpy(),o)
del ind, vals; torch.cuda.empty_cache()
def timed(fn):
    def run(n):
        oo=S.default_opts(); oo[3],oo[1],oo[4]=n,0.0,0
        t0=time.perf_counter(); fit=fn(oo)[0]; return time.perf_counter()-t0, fit
    run(1)
    (ta,f),(tb,_)=run(12),run(2)
    return (ta-tb)/10*1e3, f
for devs in ([0],[0,1]):
    for part in ("1","0"):
        if len(devs)==1 and part=="0": continue
        os.environ["SPLATT_B200_PARTITIONED_TAIL"]=part
        mg=S.MultiGpu(csf.ptr,A.CSF_TWOMODE,R,devs)
        ms,fit=timed(lambda oo: mg.cpd_als(oo,seed=1))
        print(f"devices {devs} partitioned={part} multicast={mg.multicast}: {ms:.2f} ms/iteration fit {fit:.3e}", flush=True)
        mg.free()
PY
timeout 900 python /tmp/cpd5.py 2>&1 | tail -5
echo "=== test_fused N=2"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/test_fused.py > gpurun_out/r2/fused2.log 2>&1; echo "fused rc=$?"
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/r2/fused2.log | tail -5
echo "=== bench N=2"
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5) > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err; echo "bench rc=$?"
tail -4 gpurun_out/r2/bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/bench_n2.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","parity_rel_fro","step_ms_min","step_ms_max"):
        print(k, d.get(k))
    e=d["e2e"]; print("e2e", {k:e.get(k) for k in ("value","ms_per_step","pinned","pageable_over_pinned","error")})
    for k,v in (d.get("named_configs") or {}).items():
        print("named",k, {kk:v.get(kk) for kk in ("ms_per_step","per_mode_ms","parity_rel_fro","error")})
    print("cpd", d["cpd_als_iteration"])
except Exception as e:
    print("parse failed", e)
    print(open("gpurun_out/r2/bench_n2.json").read()[-3000:])
PY
