# all BASELINE configurations through the public API + the extra lines; long series; script-recipe sweeps (1 GPU)
python bench_configs.py 1 2 4 5 6 7 > gpurun_out/configs_r2.jsonl 2> gpurun_out/configs_r2.err; tail -c 300 gpurun_out/configs_r2.err
python -m bayesflare_b200.longseries --samples 130000 2>/dev/null | tail -1 > gpurun_out/longseries_r2.jsonl
python -m bayesflare_b200.longseries --samples 1040000 2>/dev/null | tail -1 >> gpurun_out/longseries_r2.jsonl
python -m bayesflare_b200.sweep --realisations 10000 --device-sim --recipe script --sinusoid-amp 3.0 2>/dev/null | tail -1 > gpurun_out/sweep_r2.jsonl
python -m bayesflare_b200.sweep --realisations 10000 --device-sim --injtransit --sinusoid-amp 3.0 2>/dev/null | tail -1 >> gpurun_out/sweep_r2.jsonl
python -m bayesflare_b200.sweep --realisations 100000 --device-sim --recipe script 2>/dev/null | tail -1 >> gpurun_out/sweep_r2.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/configs_r2.jsonl'):
    d=json.loads(l); print(d['config'], {k:(round(v,6) if isinstance(v,float) else v) for k,v in d.items() if k in ('units_per_s_api','units_per_s_abi','api_steady_s','abi_s','abi_batch_s','parity_max_scaled_err','units_per_s_window_kernels','api_batch_s','all_device_sweep_s','transit_parity_max_scaled_err')})
for f in ('longseries_r2','sweep_r2'):
    for l in open('gpurun_out/%s.jsonl'%f):
        d=json.loads(l); print(f, {k:v for k,v in d.items() if k not in ('injected','detected','what')})
PY
