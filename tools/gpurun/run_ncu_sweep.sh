ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/sweep_launches.csv python -m bayesflare_b200.sweep --realisations 4096 --device-sim > gpurun_out/ncu_sweep.log 2>&1
python - <<'PY'
import csv, collections, re
rows=list(csv.reader(l for l in open('gpurun_out/sweep_launches.csv') if not l.startswith('==')))
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); iu=h.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    if len(r)<=iv: continue
    v=float(r[iv].replace(',','')); u=r[iu]
    ms = v/1e6 if u in ('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
    k=re.sub(r'\(.*','',r[ik]).strip()[:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=ms
for k,(n,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print("%-62s %4d %9.3f ms"%(k,n,ms))
PY
