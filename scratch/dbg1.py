import sys; sys.path.insert(0,'.')
import numpy as np
import bayesflare_b200 as bf
from bayesflare_b200 import _lib, bayes as bb
from oracle import oracle as orc
here = np.load('scratch/bg55_here.npy')
box = bb.background_models(55,4)
print('bgmodels differ from container numpy at', np.argwhere(here!=box).tolist())
try:
    from numpy._core._multiarray_umath import __cpu_features__ as f
    print('AVX512F', f.get('AVX512F'), 'AVX512_SKX', f.get('AVX512_SKX'))
except Exception as e: print(e)
g = np.load('tests/golden/kat200.npz')
ctx = _lib.default_context()
clc, cle, cts = g['clc'], g['cle'], g['cts']
sk=0.9
for bgm, nm in ((box,'box numpy'),(here,'container numpy')):
    plan = _lib.Plan(ctx, 55, bgm, cts[:55], [])
    v = sk**2+cle**2
    sls = float(np.sum(np.log(np.sqrt(v))))
    _, B = plan.window_lnB(0, clc, cle, sk, sls, want_lnB=False, want_bg=True)
    ref = g['ref_bgonly']
    o = orc.background_only(clc, cle, cts, sk)
    print(nm, 'gpu-ref first 6', np.abs(B-ref)[:6], 'max', np.abs(B-ref).max(), 'oracle-ref max', np.abs(o-ref).max(), 'gpu-oracle first 6', np.abs(B-o)[:6])
