# per-launch device times of one step at a given batch size (ncu launch list; cold-cache, serialised)
B=${1:-128}
NO_ORACLE=1 ITERS=3 python tools/time_stage.py $B 4400 > gpurun_out/plain_list.log 2>&1 &&
NO_ORACLE=1 ITERS=3 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/list_$B.csv python tools/time_stage.py $B 4400 > gpurun_out/ncu_list.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/list_$B.csv') if not l.startswith('=='))]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
for r in rows[-16:]:
    if len(r)>iv: print("%-50s %10s %s"%(r[ik][:50], r[iv], r[iu]))
PY
