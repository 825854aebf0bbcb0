python - <<'PY'
import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
from bayesflare_b200 import fitsio
g = np.load("tests/golden/ingest.npz")
os.makedirs("gpurun_out/fits", exist_ok=True)
rng = np.random.default_rng(0)
for i in range(40):
    flux = g["flux32"].copy()
    if i % 3: flux = np.float32(12000.) + (flux - np.float32(12000.)) * np.float32(0.02) + rng.standard_normal(len(flux)).astype(np.float32) * np.float32(8.)
    fitsio.write_lightcurve("gpurun_out/fits/kplr%09d_llc.fits" % (1000 + i), g["time_days"], flux, g["err32"], kepid=1000 + i)
PY
python -m bayesflare_b200.search 'gpurun_out/fits/*.fits' --out gpurun_out/search1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 -m bayesflare_b200.search 'gpurun_out/fits/*.fits' --out gpurun_out/search2.json 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tail -3
python - <<'PY'
import json
a=json.load(open("gpurun_out/search1.json")); b=json.load(open("gpurun_out/search2.json"))
for k in ("Star list","Flaring stars","No. flares","No. stars analysed"): assert a[k]==b[k], k
assert a["Flaring star data"]==b["Flaring star data"]
print("1-GPU and 2-GPU searches identical:", a["No. stars analysed"], "stars,", a["No. flares"], "flares in", len(a["Flaring stars"]), "stars")
PY
rm -rf gpurun_out/fits
