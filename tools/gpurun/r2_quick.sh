python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sustained > gpurun_out/q_n1.json 2> gpurun_out/q_n1.err; tail -c 300 gpurun_out/q_n1.err
python bench.py --steps 20 --warmup 5 --curves 128 --no-cpu-baseline > gpurun_out/q_c128.json 2> gpurun_out/q_c128.err; tail -c 300 gpurun_out/q_c128.err
python - <<'PY'
import json
for f in ("q_n1","q_c128"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
    except Exception as e:
        print(f, "no json", e); continue
    print(f, "value %.3e ms/step %.4f e2e %.3e e2e_search %.3e launches %d kernel_ms %.4f serial %.4f sustained %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_search"]["value"], d["gpu_launches"], d["roofline"]["kernel_ms"], d["roofline"]["serial_ms_per_step"], d.get("sustained",{}).get("ms_per_step")))
PY
