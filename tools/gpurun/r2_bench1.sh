python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 1200 gpurun_out/bench_r2_n1.err | grep -v Warning | tail -8
python bench.py --steps 20 --warmup 5 --curves 128 --no-cpu-baseline > gpurun_out/bench_r2_c128.json 2> gpurun_out/bench_r2_c128.err; tail -c 600 gpurun_out/bench_r2_c128.err
python - <<'PY'
import json
for f in ("bench_r2_n1","bench_r2_c128"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
    except Exception as e:
        print(f, "no json", e); continue
    print(f, "value %.3e ms/step %.4f e2e %.3e e2e_search %.3e launches %d kernel_ms %.4f serial %.4f frac %.3f exec %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_search"]["value"], d["gpu_launches"], d["roofline"]["kernel_ms"], d["roofline"]["serial_ms_per_step"], d["roofline"]["frac"], d["roofline"]["executed"]["frac"]))
    if "sustained" in d: print("   sustained", {k:v for k,v in d["sustained"].items() if k not in ("clocks","probe_clocks")})
    for k in ("cpu_baseline","parity_vs_oracle","detections"):
        if k in d: print("   ",k, d[k])
PY
