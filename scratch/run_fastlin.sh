python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -4
python bench_configs.py 1 2 4 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['config'], {k:v for k,v in d.items() if k in ('api_steady_s','abi_s','parity_max_scaled_err','detections_equal','argmax_equal_10x10')})"
BFL_NO_SPLIT=1 python tools/time_window.py 2>&1 | tail -1 | cut -c1-70
