python tools/time_pipe.py 128 3 300
python tools/time_pipe.py 128 4 300
python tools/time_pipe.py 128 6 300
python tools/time_pipe.py 128 3 20 > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 60 --csv --log-file gpurun_out/list_pipe.csv python tools/time_pipe.py 128 3 20 > gpurun_out/ncu_pipe.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/list_pipe.csv') if not l.startswith('=='))]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
for r in rows[-24:]:
    if len(r)>iv: print("%-60s %10s %s"%(r[ik][:60], r[iv], r[iu]))
PY
