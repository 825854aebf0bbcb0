"""CPU tests (`-m "not gpu"`): the C-ABI library loads and exports every symbol include/bfl.h
declares, fails loudly without a device, and the host-side mirror of the reference interface
(models, priors, noise estimators, grid weights, window time stamps) agrees with the oracle and
the golden fixtures.  No device compute here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden
import bayesflare_b200 as bf
from bayesflare_b200 import _lib, bayes as bbayes
from oracle import oracle as orc

DT = 1765.55929


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "bfl.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bfl_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libbfl.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().bfl_version() == 100


def test_no_silent_cpu_fallback():
    if _lib.load().bfl_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.BflError) as ei:
        _lib.Context()
    assert ei.value.code == -5
    ts = np.arange(300) * DT
    lc = bf.Lightcurve(cts=ts, clc=np.random.default_rng(0).standard_normal(300))
    with pytest.raises(_lib.BflError):
        bf.OddsRatioDetector(lc).oddsratio()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bayesflare_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
            assert "/root/reference" not in src, fn


MODELS = [
    (bf.Flare, "flare", {"t0": (np.inf,), "taugauss": (0, 5400, 5), "tauexp": (1800, 10800, 4), "amp": (1.,)}, {}),
    (bf.Flare, "flare", {"t0": (np.inf,), "taugauss": (0, 5400, 3), "tauexp": (0, 10800, 3), "amp": (2.5,)}, {"reverse": True}),
    (bf.Expdecay, "expdecay", {"t0": (np.inf,), "tauexp": (0.0, 900, 3), "amp": (1.,)}, {}),
    (bf.Expdecay, "expdecay", {"t0": (np.inf,), "tauexp": (0.0, 900, 3), "amp": (1.,)}, {"reverse": True}),
    (bf.Impulse, "impulse", {"t0": (0, 54 * DT, 55), "amp": (1.,)}, {}),
    (bf.Step, "step", {"t0": (np.inf,), "amp": (1.,)}, {}),
    (bf.Transit, "transit", {"t0": (np.inf,), "sigmag": (1800, 7200, 3), "tauf": (0, 7200, 4), "amp": (1.,)}, {}),
    (bf.Gaussian, "gaussian", {"t0": (np.inf,), "sigma": (0, 7200, 4), "amp": (1.,)}, {}),
]


@pytest.mark.parametrize("cls,mtype,pr,kw", MODELS)
def test_models_and_priors_match_oracle(cls, mtype, pr, kw):
    ts = np.arange(55) * DT + 1234.5
    m = cls(ts, paramranges=dict(pr), **kw)
    ranges, shape = orc.make_ranges(mtype, pr)
    assert list(m.shape) == shape
    G = int(np.prod(shape))
    par = m.grid_params()
    for q in range(G):
        idx = np.unravel_index(q, shape)
        pd = {p: ranges[p][idx[i]] for i, p in enumerate(orc.MODEL_PARAMNAMES[mtype])}
        np.testing.assert_array_equal(m(q, ts=ts).clc, orc.eval_model(mtype, pd, ts, kw.get("reverse", False)))
        np.testing.assert_array_equal(par[q, :len(idx)], [pd[p] for p in orc.MODEL_PARAMNAMES[mtype]])
        if mtype != "gaussian":
            a, b = m.prior(pd), orc.model_prior(mtype, pd, ranges)
            assert a == b or (np.isinf(a) and np.isinf(b))


def test_grid_logweight_equals_nested_logtrapz():
    rng = np.random.default_rng(3)
    ts = np.arange(55) * DT
    m = bf.Flare(ts, paramranges={"t0": (np.inf,), "taugauss": (0, 5400, 6), "tauexp": (1800, 10800, 7), "amp": (1.,)})
    lnB = rng.uniform(-30, 10, size=tuple(m.shape))
    lnB[0, 2, 3, 0] = -np.inf
    nested = orc.marginalise_full(lnB[..., None], m.ranges, m.paramnames)[0]
    lw = m.grid_logweight().reshape(m.shape)
    x = (lnB + lw).ravel()
    x = x[np.isfinite(x)]
    direct = x.max() + np.log(np.exp(x - x.max()).sum())
    assert abs(direct - nested) < 1e-12


def test_noise_estimators_against_reference_golden():
    g = load_golden("detector")
    lc = bf.Lightcurve(cts=g["a_cts"], clc=g["a_clc"])
    np.testing.assert_allclose(bf.savitzky_golay(lc.clc, 55, 4), g["ref_a_sgfit"], rtol=0, atol=1e-12)
    assert abs(bbayes.estimate_sigma(lc, 55, 4, "powerspectrum", 0.5, 1.0) - float(g["ref_a_sk_ps"])) < 1e-13
    assert abs(bbayes.estimate_sigma(lc, 55, 4, "tailveto", 0.5, 1.0) - float(g["ref_a_sk_tv"])) < 1e-13
    assert lc.detrended is False      # the estimate works on a copy (bayes.py:236)


def test_window_times_and_background_models():
    ts = np.arange(4400) * DT
    mts = bbayes.window_times(ts, 55, ts[2200])
    np.testing.assert_array_equal(mts, orc.window_times(ts, 55))
    assert len(mts) == 55
    np.testing.assert_array_equal(bbayes.background_models(55, 4), orc.background_models(55, 4))


def test_scalar_helpers_against_reference_golden():
    g = load_golden("scalars")
    np.testing.assert_array_equal([bf.logplus(x, y) for x, y in zip(g["x"], g["y"])], g["ref_logplus"])
    np.testing.assert_array_equal([bf.logminus(x, y) for x, y in zip(g["x"], g["y"])], g["ref_logminus"])
    off = 0
    for n, ref in zip(g["lt_sizes"], g["ref_logtrapz"]):
        got = bf.logtrapz(g["lt_lx"][off:off + n], g["lt_t"][off:off + n])
        off += n
        assert got == ref or abs(got - ref) <= 2e-15 * abs(ref)


def test_contiguous_regions_and_detector_config():
    cond = np.array([1, 1, 0, 0, 1, 0, 1, 1, 1], dtype=bool)
    np.testing.assert_array_equal(bf.contiguous_regions(cond), [[0, 2], [4, 5], [6, 9]])
    np.testing.assert_array_equal(bf.contiguous_regions(cond), orc.contiguous_regions(cond))
    ts = np.arange(300) * DT
    lc = bf.Lightcurve(cts=ts, clc=np.zeros(300))
    det = bf.OddsRatioDetector(lc)
    assert det.flareparams["amp"] == (1.,) and det.flareparams["t0"] == (np.inf,)
    models = det.hypotheses(ts)
    assert [m.mtype for m, _ in models] == ["flare", "impulse", "expdecay", "expdecay"]
    assert [hr for _, hr in models] == [True, False, True, True]
    assert sum(int(np.isfinite(m.grid_logprior()).sum()) for m, _ in models) + 1 == 93   # G_eval
    with pytest.raises(ValueError):
        det.set_flare_params({"tauexp": (1, 2, 3)})
    with pytest.raises(NotImplementedError):
        bf.OddsRatioDetector(lc, nsinusoids=2).plan()


def test_impulse_excluder_m_shaped_features():
    """find.py:866-912: a dip below -5 flanked by rising then falling positive pairs removes
    [idx - width, idx + width) (clamped); anything else is kept"""
    import bayesflare_b200 as bf
    n = 60
    lnO = np.full(n, -1.0)
    ts = np.arange(n, dtype=float)
    for c in (20, 3, 57):                       # interior M shape, one near each end (clamped windows)
        lnO[c - 2:c + 3] = [1.0, 3.0, -8.0, 3.0, 1.0]
    lnO[40] = -9.0                              # a dip without the flanks: kept
    lnO[48:53] = [3.0, 1.0, -8.0, 3.0, 1.0]     # left flank falls instead of rising: kept
    det = bf.OddsRatioDetector(bf.Lightcurve(cts=ts * 1765.5, clc=np.zeros(n)))
    got, gts = det.impulse_excluder(lnO, ts, exclusionwidth=5)
    keep = np.ones(n, dtype=bool)
    keep[15:25] = False
    keep[0:8] = False
    keep[52:59] = False                         # enidx clamps to len - 1, the last sample survives
    np.testing.assert_array_equal(gts, ts[keep])
    np.testing.assert_array_equal(got, lnO[keep])
    got2, _ = det.impulse_excluder(list(lnO), ts, exclusionwidth=2)
    assert len(got2) == n - 4 - 4 - 4
