T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/f_n1.json 2> gpurun_out/f_n1.err
$T --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --steps 20 --warmup 5 --no-weak > gpurun_out/f_n8.json 2> gpurun_out/f_n8.err; tail -c 300 gpurun_out/f_n8.err | grep -i "error\|Traceback"
$T --nproc-per-node 4 --master-port 29713 bench.py --gpus 4 --steps 20 --warmup 5 --no-weak --no-sustained > gpurun_out/f_n4.json 2> gpurun_out/f_n4.err
python tools/bench_summary.py gpurun_out/f_n1.json gpurun_out/f_n4.json gpurun_out/f_n8.json
python - <<'PY'
import json
for f in ("f_n1","f_n4","f_n8"):
    d=json.loads([l for l in open("gpurun_out/%s.json"%f) if l.startswith("{")][-1]); print(f, d["ms_per_step_repeats"], d["e2e"].get("cuda_graphs"))
PY
