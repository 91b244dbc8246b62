python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']; print(d['value']/1e12, d['e2e']['value']/1e12, r['frac'], r['gather_path']['frac_of_measured_gather_peak'], d['cpu_baseline']['value']/1e9, d['cpu_baseline']['cores'], d['cpd_als_iteration'], d['clocks'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
