python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_lin.json
python tools/time_window.py 2>&1 | tail -2
python -c "
import json; d=json.load(open('gpurun_out/bench_lin.json')); print(d['value'],d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
"
