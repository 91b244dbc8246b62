timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']; print(d['value']/1e12, d['e2e']['value']/1e12, r['frac'], r['gather_path'].get('frac_of_measured_gather_peak'), d['cpu_baseline']['value']/1e9, d['cpu_baseline']['cores'], d['cpd_als_iteration'], d['clocks'], d['gpu_launches'])"
