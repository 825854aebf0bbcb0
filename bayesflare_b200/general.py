"""Log-space scalar helpers with the reference's names and semantics
(``bayesflare/stats/general.pyx:519-607``).  Scalar / 1-D convenience functions for user code
and post-processing; the bulk marginalisation over model grids runs on the GPU
(``bfl_logtrapz`` and the fused log-sum-exp in the window kernel)."""
from __future__ import annotations

from math import exp, isinf, log

import numpy as np

__all__ = ["logplus", "logminus", "logtrapz", "LOG2"]

LOG2 = 0.693147180559945


def logplus(x, y):
    """log(e^x + e^y)  (general.pyx:519-547)"""
    x, y = float(x), float(y)
    z = np.inf
    bothinf = isinf(x) and isinf(y)
    if bothinf and x < 0 and y < 0:
        z = -np.inf
    elif x > y and not bothinf:
        z = x + log(1 + exp(y - x))
    elif x <= y and not bothinf:
        z = y + log(1 + exp(x - y))
    return z


def logminus(x, y):
    """log(e^x - e^y)  (general.pyx:550-573, including its ``y < x -> -inf`` branch order)"""
    x, y = float(x), float(y)
    z = np.inf
    bothinf = isinf(x) and isinf(y)
    if (bothinf and x < 0 and y < 0) or y < x:
        z = -np.inf
    elif x > y and not bothinf:
        z = x + log(1 - exp(y - x))
    return z


def logtrapz(lx, t):
    """log of the trapezium-rule integral of exp(lx) over t  (general.pyx:576-607)"""
    lx = np.asarray(lx)
    t = np.asarray(t)
    assert lx.dtype == np.float64 and t.dtype == np.float64
    with np.errstate(divide="ignore", invalid="ignore"):
        ldts = np.log(np.diff(t))
    B = -np.inf
    for i in range(len(lx) - 1):
        z = logplus(lx[i], lx[i + 1])
        z = z + ldts[i] - LOG2
        B = logplus(z, B)
    return B
