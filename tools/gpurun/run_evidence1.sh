python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_err.log; tail -c 300 gpurun_out/bench_err.log
python bench_configs.py 1 2 4 5 6 > gpurun_out/configs_r1.jsonl 2>gpurun_out/configs_err.log; tail -c 300 gpurun_out/configs_err.log
python -c "
import json
d=json.load(open('gpurun_out/bench_r1_n1.json')); print(d['value'], d['ms_per_step'], d['steps'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['executed']['frac'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
for l in open('gpurun_out/configs_r1.jsonl'):
    d=json.loads(l); print(d['config'], {k:v for k,v in d.items() if k in ('units_per_s_api','units_per_s_abi','api_steady_s','abi_s','abi_batch_s','parity_max_scaled_err','units_per_s_window_kernels','api_batch_s','all_device_sweep_s')})
"
