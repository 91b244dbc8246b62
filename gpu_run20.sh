python bench.py --steps 50 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']; print(d['value']/1e12, d['e2e']['value']/1e12, r['frac'], r['gather_path'], d['cpu_baseline'], d['cpd_als_iteration'], d['clocks'])"
