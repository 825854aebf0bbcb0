T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sustained > gpurun_out/bench_r2_n1b.json 2> gpurun_out/bench_r2_n1b.err
$T --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-weak > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; tail -c 300 gpurun_out/bench_r2_n8.err
python - <<'PY'
import json
base=None
for f in ("bench_r2_n1b","bench_r2_n8"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no json", e); continue
    if base is None: base=d
    n=d["n_gpus"]
    print(f, "N=%d value %.3e (eff %.3f) ms/step %.4f e2e %.3e (eff %.3f) e2e_search %.3e (eff %.3f) coll %s sustained %s" % (n, d["value"], d["value"]/base["value"]/n, d["ms_per_step"], d["e2e"]["value"], d["e2e"]["value"]/base["e2e"]["value"]/n, d["e2e_search"]["value"], d["e2e_search"]["value"]/base["e2e_search"]["value"]/n, d["collective"]["us"], d.get("sustained",{}).get("ms_per_step")))
PY
