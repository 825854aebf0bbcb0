import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import bayesflare_b200 as bf
from test_gpu_parity import _varying_cle_case, _lc
ts, clc, cle = _varying_cle_case(600, 9)
clc = clc.copy(); clc[300] = np.nan
det = bf.OddsRatioDetector(_lc(ts, clc, cle))
got = np.ravel(det.plan().detector_odds(clc, cle, 1.0, k0=27, k1=600 - 27))
bad = np.nonzero(~np.isfinite(got))[0]
print(bad.min(), bad.max(), len(bad), got[244:249], got[299:303])
got = np.ravel(det.plan().detector_odds(clc, None, 1.0, k0=27, k1=600 - 27))
bad = np.nonzero(~np.isfinite(got))[0]
print("const", bad.min(), bad.max(), len(bad))
