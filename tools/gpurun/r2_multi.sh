# 8-GPU box: strong-scaling curve of bench.py, long series as time chunks, injection sweeps (script recipe)
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
bash tools/gpurun/r2_scale.sh 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12
: > gpurun_out/longseries_multi.jsonl
for n in 2 4 8; do
  $T --nproc-per-node $n --master-port $((29600+n)) -m bayesflare_b200.longseries --samples 1040000 2>/dev/null | grep "^{" >> gpurun_out/longseries_multi.jsonl
done
$T --nproc-per-node 8 --master-port 29620 -m bayesflare_b200.longseries --samples 8320000 2>/dev/null | grep "^{" >> gpurun_out/longseries_multi.jsonl
python -m bayesflare_b200.longseries --samples 8320000 2>/dev/null | grep "^{" >> gpurun_out/longseries_multi.jsonl
: > gpurun_out/sweep_multi.jsonl
for n in 1 2 4 8; do
  if [ $n = 1 ]; then R="python"; else R="$T --nproc-per-node $n --master-port $((29640+n))"; fi
  $R -m bayesflare_b200.sweep --realisations 100000 --device-sim --recipe script --sinusoid-amp 3.0 2>/dev/null | grep "^{" >> gpurun_out/sweep_multi.jsonl
  $R -m bayesflare_b200.sweep --realisations 100000 --device-sim --injtransit --sinusoid-amp 3.0 2>/dev/null | grep "^{" >> gpurun_out/sweep_multi.jsonl
done
python - <<'PY'
import json
for f in ("longseries_multi","sweep_multi"):
    for l in open("gpurun_out/%s.jsonl"%f):
        d=json.loads(l); print(f, {k:(round(v,5) if isinstance(v,float) else v) for k,v in d.items() if k not in ('injected','detected','what','units','t_simulate_s_rank0')})
PY
