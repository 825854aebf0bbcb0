"""Print the headline numbers of bench.py JSON lines; the first file is the N = 1 reference for the efficiencies.
    python tools/bench_summary.py a.json b.json ..."""
import json, sys
base = None
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no json line:", e)
        continue
    if base is None:
        base = d
    n = d["n_gpus"]
    def eff(a, b):
        return a / b / n * base["n_gpus"] if b else float("nan")
    sus, bs = d.get("sustained", {}), base.get("sustained", {})
    full, bfull = d.get("e2e_full_lnO", {}), base.get("e2e_full_lnO", {})
    print("%s: N=%d curves/gpu=%d value %.3e (eff %.3f) ms/step %.4f | e2e %.3e (eff %.3f) | e2e_full %.3e (eff %.3f) | "
          "sustained %.3e (eff %.3f) | kernel_ms %.4f frac %.3f exec %.3f | launches %d coll_us %s weak %s"
          % (f.split("/")[-1], n, d["config"]["curves_per_gpu"], d["value"], eff(d["value"], base["value"]), d["ms_per_step"],
             d["e2e"]["value"], eff(d["e2e"]["value"], base["e2e"]["value"]),
             full.get("value", 0), eff(full.get("value", 0), bfull.get("value", 0)),
             sus.get("value", 0), eff(sus.get("value", 0), bs.get("value", 0)),
             d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["executed"]["frac"],
             d["gpu_launches"], d.get("collective", {}).get("us"), d.get("weak_scaling", {}).get("value")))
    for k in ("cpu_baseline", "cpu_baseline_port", "parity_vs_oracle"):
        if k in d:
            print("    %s: %s" % (k, d[k]))
