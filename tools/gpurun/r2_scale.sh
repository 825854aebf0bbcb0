# N GPUs of one box: the strong-scaling curve (bench.py at 1, 2, 4, 8 ranks) and the PCIe aggregate
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NS=${NS:-"2 4 8"}
python bench.py --steps 20 --warmup 5 > gpurun_out/sc_n1.json 2> gpurun_out/sc_n1.err
FILES=gpurun_out/sc_n1.json
P=29520
for n in $NS; do
  P=$((P+1))
  $T --nproc-per-node $n --master-port $P bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/sc_n$n.json 2> gpurun_out/sc_n$n.err; tail -c 200 gpurun_out/sc_n$n.err | grep -i "error\|Traceback"
  FILES="$FILES gpurun_out/sc_n$n.json"
done
python tools/bench_summary.py $FILES
