"""GPU parity tests (`-m gpu`): the CUDA path, called through the C ABI / the drop-in Python API,
against (a) golden fixtures generated from the reference and (b) the CPU oracle on seeded inputs.

Tolerance (BASELINE.json north_star): |delta| <= 1e-9 * max(1, |ref|) on log odds / log Bayes
factors with an identical -inf pattern; detection indices bit-exact."""
import os

import numpy as np
import pytest

from conftest import assert_logodds_close, load_golden
import bayesflare_b200 as bf
from bayesflare_b200 import _lib, bayes as bbayes
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
DT = 1765.55929
TOL = 1e-9


def _lc(cts, clc, cle=None):
    return bf.Lightcurve(cts=cts, clc=clc, cle=cle)


def _force_sk(monkeypatch, sk):
    monkeypatch.setattr(bbayes, "estimate_sigma", lambda *a, **k: sk)


KAT = {
    "flare": (bf.Flare, {}, {"t0": (np.inf,), "taugauss": (0, 5400, 4), "tauexp": (1800, 10800, 4), "amp": (1.,)}, True),
    "impulse": (bf.Impulse, {}, {"t0": (0, 54 * DT, 55), "amp": (1.,)}, False),
    "impulse1": (bf.Impulse, {}, {"t0": (np.inf,), "amp": (1.,)}, False),
    "expdecay": (bf.Expdecay, {}, {"t0": (np.inf,), "tauexp": (0.0, 900, 3), "amp": (1.,)}, True),
    "expdecay_rev": (bf.Expdecay, {"reverse": True}, {"t0": (np.inf,), "tauexp": (0.0, 900, 3), "amp": (1.,)}, True),
    "step": (bf.Step, {}, {"t0": (np.inf,), "amp": (1.,)}, False),
    "transit": (bf.Transit, {}, {"t0": (np.inf,), "sigmag": (1800, 7200, 3), "tauf": (0, 7200, 4), "amp": (1.,)}, True),
    "flare_full": (bf.Flare, {}, {"t0": (np.inf,), "taugauss": (0, 5400, 3), "tauexp": (1800, 10800, 3), "amp": (1.,)}, False),
}


def test_device_present_and_native_library_loaded():
    assert _lib.load().bfl_device_count() >= 1
    ctx = _lib.default_context()
    tf, ms = _lib.fp64_peak(ctx, 512)
    assert tf > 1.0, "FP64 probe reports %.2f TFLOP/s" % tf
    assert ctx.launches > 0


@pytest.mark.parametrize("name", sorted(KAT))
def test_templates_generated_on_device(name):
    cls, kw, pr, hr = KAT[name]
    ts = np.arange(400) * DT
    m = cls(ts, paramranges=dict(pr), **kw)
    mts = bbayes.window_times(ts, 55, m.t0)
    plan = _lib.Plan(_lib.default_context(), 55, bbayes.background_models(55, 4), mts,
                     [bbayes.hypothesis_from_model(m, mts, hr)])
    dev = plan.templates(0)
    lp = m.grid_logprior()
    for q in range(dev.shape[0]):
        host = m(q, ts=mts).clc if np.isfinite(lp[q]) else np.zeros(55)
        np.testing.assert_allclose(dev[q], host, rtol=4e-16, atol=1e-300, err_msg="%s q=%d" % (name, q))
        assert np.array_equal(dev[q] == 0, host == 0)


@pytest.mark.parametrize("name", sorted(KAT))
def test_kat200_general_variance_path_vs_reference(name, monkeypatch):
    """Per-sample variance (cle varies): every sample goes through the reference-order path."""
    g = load_golden("kat200")
    _force_sk(monkeypatch, float(g["sk"]))
    cls, kw, pr, hr = KAT[name]
    lc = _lc(g["cts"], g["clc"], g["cle"])
    B = bf.Bayes(lc, cls(lc.cts, amp=1, paramranges=dict(pr), **kw))
    B.bayes_factors_marg_poly_bgd(halfrange=hr)
    assert_logodds_close(B.lnBmargAmp, g["ref_lnB_" + name], TOL, name)
    assert_logodds_close(B.marginalise_full().lnBmargAmp, g["ref_marg_" + name], TOL, name + " marg")
    if name == "flare":
        assert_logodds_close(B.bayes_factors_marg_poly_bgd_only(), g["ref_bgonly"], TOL, "bgonly")
        assert abs(B.noise_ev - float(g["ref_noise_ev"])) <= 1e-9 * abs(float(g["ref_noise_ev"]))
        assert abs(B.lnBmargAmp[0, 3, 0, 0, 100] - (-19.112936560497506)) < 1e-9


def test_kat200_other_window_and_order(monkeypatch):
    g = load_golden("kat200")
    _force_sk(monkeypatch, float(g["sk"]))
    lc = _lc(g["cts"], g["clc"], g["cle"])
    B = bf.Bayes(lc, bf.Flare(lc.cts, amp=1, paramranges=dict(KAT["flare"][2])))
    B.bayes_factors_marg_poly_bgd(bglen=21, bgorder=2)
    assert_logodds_close(B.lnBmargAmp, g["ref_lnB_flare_w21_o2"], TOL, "w21 o2")
    assert_logodds_close(B.bayes_factors_marg_poly_bgd_only(bglen=21, bgorder=2), g["ref_bgonly_w21_o2"], TOL, "bg w21")


@pytest.mark.parametrize("name", ["flare", "impulse", "expdecay_rev", "step", "transit", "flare_full"])
@pytest.mark.parametrize("bglen,bgorder", [(55, 4), (21, 2), (55, 0), (33, 5)])
def test_constant_variance_fast_path_vs_oracle(name, bglen, bgorder, monkeypatch):
    """cle constant: interior samples take the translation-invariant fast path, the first/last
    bglen/2 samples the truncated-window path; all N samples are compared."""
    rng = np.random.default_rng(42)
    N = 700
    ts = np.arange(N) * DT
    clc = rng.standard_normal(N) * 1.3 + orc.flare_model(ts, ts[300], 9., 1500., 4000.) \
        + 2.0 * np.sin(2 * np.pi * ts / (2.2 * 86400.))
    clc -= np.median(clc)
    cle = np.full(N, 0.4)
    sk = 1.234
    _force_sk(monkeypatch, sk)
    cls, kw, pr, hr = KAT[name]
    pr = dict(pr)
    if name == "impulse":
        pr["t0"] = (0, (bglen - 1) * DT, bglen)
    lc = _lc(ts, clc, cle)
    B = bf.Bayes(lc, cls(ts, amp=1, paramranges=pr, **kw))
    B.bayes_factors_marg_poly_bgd(bglen=bglen, bgorder=bgorder, halfrange=hr)
    ref, _, refbg = orc.bayes_factors(clc, cle, ts, cls(ts).mtype, pr, sk, bglen=bglen, bgorder=bgorder,
                                      halfrange=hr, reverse=kw.get("reverse", False), with_background=True)
    assert_logodds_close(B.lnBmargAmp, ref, TOL, "%s W%d o%d" % (name, bglen, bgorder))
    assert_logodds_close(B.bayes_factors_marg_poly_bgd_only(bglen=bglen, bgorder=bgorder), refbg, TOL, "bg")


@pytest.mark.parametrize("meth", ["powerspectrum", "tailveto"])
def test_detector_default_vs_reference(meth):
    g = load_golden("detector")
    det = bf.OddsRatioDetector(_lc(g["a_cts"], g["a_clc"]), noiseestmethod=meth)
    lnO, ts = det.oddsratio()
    assert isinstance(lnO, list) and len(lnO) == len(g["a_clc"]) - 54
    assert_logodds_close(lnO, g["ref_a_lnO_" + meth], TOL, meth)
    np.testing.assert_array_equal(ts, g["a_cts"][27:-27])
    fl, n, mx = det.thresholder(lnO, 16.5, expand=1)
    assert np.array_equal(np.array(fl, dtype=np.int64).reshape(-1, 2), g["ref_a_segs_" + meth])
    assert np.array_equal(np.array([m[1] for m in mx], dtype=float), g["ref_a_max_" + meth][:, 1])
    assert det.ngrid_eval == 93 and det.ngrid_nominal == 108


def test_detector_scripts_configuration_and_options_vs_reference():
    g = load_golden("detector")
    det = bf.OddsRatioDetector(_lc(g["b_cts"], g["b_clc"]), bglen=55, bgorder=4, noiseestmethod="tailveto", tvsigma=1.0,
                               noiseimpulseparams={"t0": (0, 54 * DT, 55)})
    lnO, ts = det.oddsratio()
    assert_logodds_close(lnO, g["ref_b_lnO"], TOL, "scripts detector")
    fl, n, mx = det.thresholder(lnO, 10.0, expand=2)
    assert np.array_equal(np.array(fl, dtype=np.int64).reshape(-1, 2), g["ref_b_segs"])
    assert np.array_equal(np.array([m[1] for m in mx], dtype=float), g["ref_b_max"][:, 1])
    assert det.ngrid_eval == 147
    det = bf.OddsRatioDetector(_lc(g["c_cts"], g["c_clc"]), noiseestmethod="tailveto", noisestep=True, ignoreedges=False)
    det.set_noise_impulse(noiseimpulse=True, noiseimpulseparams={"t0": (np.inf,)}, positive=True)
    lnO, ts = det.oddsratio()
    assert len(lnO) == len(g["c_clc"])
    assert_logodds_close(lnO, g["ref_c_lnO"], TOL, "step + edges + positive impulse")


def test_config1_single_long_cadence_quarter_vs_reference():
    """BASELINE config 1: N=4400, default detector; SURVEY 8(c) end-to-end known answer."""
    g = load_golden("config1")
    for meth, kat in (("powerspectrum", 18.343102610670442), ("tailveto", 19.80097630627027)):
        det = bf.OddsRatioDetector(_lc(g["cts"], g["clc"]), noiseestmethod=meth)
        lnO, ts = det.oddsratio()
        assert len(lnO) == 4346
        assert_logodds_close(lnO, g["ref_lnO_" + meth], TOL, "config1 " + meth)
        assert int(np.argmax(lnO)) == 2174 and abs(max(lnO) - kat) < 1e-9
        fl, n, mx = det.thresholder(lnO, 16.5, expand=1)
        assert [tuple(s) for s in fl] == [(2171, 2176)] and n == 1 and mx[0][1] == 2174
        above = np.nonzero(np.array(lnO) > 16.5)[0]
        np.testing.assert_array_equal(above, np.nonzero(g["ref_lnO_" + meth] > 16.5)[0])   # bit-exact detections


def test_config2_short_cadence_vs_reference():
    g = load_golden("shortcad")
    det = bf.OddsRatioDetector(_lc(g["cts"], g["clc"]))
    lnO, ts = det.oddsratio()
    assert_logodds_close(lnO, g["ref_lnO"], TOL, "short cadence")
    for thr in (5.0, 16.5):
        np.testing.assert_array_equal(np.nonzero(np.array(lnO) > thr)[0], np.nonzero(g["ref_lnO"] > thr)[0])


def test_detector_no_noise_models_and_varying_errors_vs_oracle():
    rng = np.random.default_rng(5)
    N = 500
    ts = np.arange(N) * DT
    clc = rng.standard_normal(N) + orc.flare_model(ts, ts[250], 7., 900., 3500.)
    cle = 0.2 + 0.1 * rng.random(N)
    det = bf.OddsRatioDetector(_lc(ts, clc, cle), noiseestmethod="tailveto", noisepoly=False, noiseimpulse=False,
                               noiseexpdecay=False)
    lnO, _ = det.oddsratio()
    ref, _ = orc.oddsratio(clc, cle, ts, noiseestmethod="tailveto", noisepoly=False, noiseimpulse=False, noiseexpdecay=False)
    assert_logodds_close(lnO, ref, TOL, "no noise models")
    det = bf.OddsRatioDetector(_lc(ts, clc, cle), noiseestmethod="tailveto", ignoreedges=False)
    lnO, _ = det.oddsratio()
    ref, _ = orc.oddsratio(clc, cle, ts, noiseestmethod="tailveto", ignoreedges=False)
    assert_logodds_close(lnO, ref, TOL, "varying cle detector")


def test_batch_equals_single_and_oracle():
    rng = np.random.default_rng(9)
    B, N = 6, 900
    ts = np.arange(N) * DT
    clc = rng.standard_normal((B, N)) * rng.uniform(0.5, 3, (B, 1))
    for b in range(B):
        clc[b] += orc.flare_model(ts, ts[100 + 120 * b], 5. + 3 * b, 800., 2500.)
    det = bf.OddsRatioDetector(_lc(ts, clc[0]), noiseestmethod="tailveto")
    sk = np.array([det.noise_sigma(_lc(ts, clc[b])) for b in range(B)])
    batch = det.oddsratio_batch(clc, sk=sk)
    assert batch.shape == (B, N - 54)
    for b in range(B):
        single, _ = bf.OddsRatioDetector(_lc(ts, clc[b]), noiseestmethod="tailveto").oddsratio()
        # tiling (samples per thread, template chunks) depends on the batch size, so the
        # log-sum-exp merge order may differ: identical to rounding, not bit for bit
        np.testing.assert_allclose(batch[b], single, rtol=0, atol=2e-12)
        ref, _ = orc.oddsratio(clc[b], np.zeros(N), ts, noiseestmethod="tailveto")
        assert_logodds_close(batch[b], ref, TOL, "batch %d" % b)


def test_time_chunks_with_halos_equal_whole_series():
    """Config-2 mechanism: a long series cut into chunks, each staged with a W/2 halo, gives
    bit-identical log odds to the un-chunked run (device-resident API)."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(11)
    N = 5000
    ts = np.arange(N) * 58.84876
    clc = rng.standard_normal(N) + orc.flare_model(ts, ts[2600], 9., 200., 900.)
    det = bf.OddsRatioDetector(_lc(ts, clc))
    plan = det.plan()
    sk = det.noise_sigma()
    whole = plan.detector_odds(clc, None, sk)
    dev = torch.device("cuda:0")
    d_sk = torch.tensor([sk], dtype=torch.float64, device=dev)
    out = np.empty(N)
    H = 27
    bounds = [0, 1300, 2617, 3999, N]
    for k0, k1 in zip(bounds[:-1], bounds[1:]):
        s0, s1 = max(0, k0 - H), min(N, k1 + H)
        d_seg = torch.tensor(clc[s0:s1], dtype=torch.float64, device=dev)
        d_out = torch.empty(k1 - k0, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        plan.detector_odds_dev(d_seg.data_ptr(), 0, True, d_sk.data_ptr(), 1, N, s0, s1 - s0, k0, k1, d_out.data_ptr())
        plan.ctx.sync()
        out[k0:k1] = d_out.cpu().numpy()
    fin = np.isfinite(whole)
    assert np.array_equal(np.isfinite(out), fin)
    np.testing.assert_allclose(out[fin], whole[fin], rtol=0, atol=2e-13)
    ref, _ = orc.oddsratio(clc, np.zeros(N), ts, ignoreedges=False)
    assert_logodds_close(out[27:-27], ref[27:-27], TOL, "chunked short cadence")
    np.testing.assert_array_equal(np.nonzero(out[27:-27] > 16.5)[0], np.nonzero(ref[27:-27] > 16.5)[0])


def test_threshold_on_device_vs_oracle_random():
    rng = np.random.default_rng(13)
    ctx = _lib.default_context()
    for n, thr, expand in [(50, 0.5, 0), (1000, 1.0, 0), (1000, 1.0, 3), (5000, 2.0, 1), (300, -9.0, 2), (300, 9.0, 2),
                           (4000, 1.5, 7)]:
        x = np.cumsum(rng.standard_normal(n)) * 0.3 + rng.standard_normal(n)
        segs, mv, mi = _lib.threshold(ctx, x, thr, expand)
        rs, rn, rm = orc.thresholder(x, thr, expand=expand)
        assert [tuple(map(int, s)) for s in segs] == [tuple(map(int, s)) for s in rs]
        assert [int(i) for i in mi] == [int(m[1]) for m in rm]
        assert [float(v) for v in mv] == [float(m[0]) for m in rm]


def test_logtrapz_on_device():
    rng = np.random.default_rng(17)
    lx = rng.uniform(-40, 5, (9, 37))
    lx[3, 5] = -np.inf
    lx[:, 7] = -np.inf
    t = np.sort(rng.uniform(0, 10, 9))
    got = _lib.logtrapz_axis0(_lib.default_context(), lx, t)
    ref = np.array([orc.logtrapz(lx[:, j], t) for j in range(37)])
    assert_logodds_close(got, ref, 1e-13, "logtrapz")


def test_parameter_estimation_grid_vs_reference():
    """BASELINE config 4 (reduced grid): scripts/parameter_estimation_example.py."""
    g = load_golden("pegrid")
    lc = _lc(g["cts"], g["clc"])
    t0, idx, a = float(g["t0"]), int(g["idxt0"]), g["amprange"]
    for bgorder in (3, 4, 2):
        pe = bf.ParameterEstimationGrid('flare', lc)
        assert abs(pe.noiseSigma - float(g["ref_sigma"])) < 1e-12
        pe.lightcurve_chunk(idx, 55)
        pe.set_grid(ranges={'taugauss': (0., 3600., 12), 'tauexp': (0., 12000., 10), 'amp': (a[0], a[1], int(a[2])), 't0': (t0,)})
        pe.calculate_posterior(bgorder=bgorder)
        assert pe.posterior.shape == (1, 10, 12, 9)
        assert_logodds_close(pe.posterior, g["ref_post_o%d" % bgorder], TOL, "pe bgorder %d" % bgorder)
        if bgorder == 3:
            pe.marginalise_all()
            for p in ("taugauss", "tauexp", "amp"):
                np.testing.assert_allclose(pe.margposteriors[p], g["ref_margpost_" + p], rtol=1e-8, atol=1e-300)
            mp, mpar = pe.maximum_posterior()
            assert abs(mp - float(g["ref_maxpost"])) <= 1e-9 * abs(float(g["ref_maxpost"]))
            np.testing.assert_array_equal([mpar[p] for p in ("t0", "tauexp", "taugauss", "amp")], g["ref_maxpost_params"])
            assert abs(pe.maximum_posterior_snr() - float(g["ref_snr"])) < 1e-9
    pe = bf.ParameterEstimationGrid('flare', lc)
    pe.lightcurve_chunk(idx, 55)
    pe.set_grid(ranges={'taugauss': (0., 3600., 6), 'tauexp': (0., 12000., 5), 'amp': (a[0], a[1], int(a[2])), 't0': (t0,)})
    pe.calculate_posterior(margbackground=False)
    assert_logodds_close(pe.posterior, g["ref_post_nobg"], TOL, "pe no background")
    # Savitzky-Golay detrended chunk: every model is filtered the same way (bayes.py:1018-1021)
    pe = bf.ParameterEstimationGrid('flare', lc)
    pe.lightcurve_chunk(idx, 55)
    pe.lightcurve.detrend(method="savitzkygolay", nbins=21, order=3)
    np.testing.assert_allclose(pe.lightcurve.clc, g["ref_sg_chunk_clc"], rtol=0, atol=1e-12)
    pe.set_grid(ranges={'taugauss': (0., 3600., 6), 'tauexp': (0., 12000., 5), 'amp': (a[0], a[1], int(a[2])), 't0': (t0,)})
    pe.calculate_posterior(bgorder=2)
    assert_logodds_close(pe.posterior, g["ref_post_sg21_o2"], TOL, "pe on a detrended chunk")


def test_dense_grid_property_and_oracle_subset():
    """BASELINE config 4 detector part: 64x64 tau grid (templates span many chunks).  Checked by
    (i) the oracle on a short curve, (ii) refinement consistency at full length."""
    rng = np.random.default_rng(21)
    N = 260
    ts = np.arange(N) * DT
    clc = rng.standard_normal(N) + orc.flare_model(ts, ts[130], 10., 1200., 3000.)
    fp = {"taugauss": (0, 5400, 64), "tauexp": (1800, 10800, 64)}
    det = bf.OddsRatioDetector(_lc(ts, clc), noiseestmethod="tailveto", flareparams=fp)
    lnO, _ = det.oddsratio()
    ref, _ = orc.oddsratio(clc, np.zeros(N), ts, noiseestmethod="tailveto", flareparams=fp)
    assert_logodds_close(lnO, ref, TOL, "64x64 grid")


def test_error_conventions(capsys):
    ts = np.arange(300) * DT
    lc = _lc(ts, np.random.default_rng(0).standard_normal(300))
    B = bf.Bayes(lc, bf.Flare(ts, paramranges={"t0": (np.inf,), "taugauss": (0, 5400, 2), "tauexp": (1800, 10800, 2), "amp": (1.,)}))
    assert B.bayes_factors_marg_poly_bgd(bglen=54) is None            # bayes.py:229-233
    assert "must be an odd number" in capsys.readouterr().out
    assert B.bayes_factors_marg_poly_bgd(noiseestmethod="nope") is None   # bayes.py:243-245
    assert "Noise estimation method" in capsys.readouterr().out
    with pytest.raises(_lib.BflError):
        _lib.Plan(_lib.default_context(), 54, bbayes.background_models(54, 4), ts[:54], [])
    # non-finite data -> -inf, as log_marg_amp_full.c:20,25 (reference-order path)
    clc = lc.clc.copy()
    clc[150] = np.nan
    m = bf.Step(ts, paramranges={"t0": (np.inf,), "amp": (1.,)})
    mts = bbayes.window_times(ts, 55, m.t0)
    plan = _lib.Plan(_lib.default_context(), 55, bbayes.background_models(55, 4), mts,
                     [bbayes.hypothesis_from_model(m, mts, False)])
    lnB, bg = plan.window_lnB(0, clc, np.linspace(0.1, 0.2, 300), 1.0, 0.0, want_bg=True)
    assert np.all(lnB[0, 150 - 27:150 + 28] == -np.inf) and np.all(bg[150 - 27:150 + 28] == -np.inf)
    assert np.all(np.isfinite(lnB[0, :150 - 27])) and np.all(np.isfinite(bg[150 + 28:]))


@pytest.mark.parametrize("meth", ["powerspectrum", "tailveto"])
def test_noise_sigma_on_device_vs_reference_and_host(meth):
    """the noise estimate on the device against the reference's value (golden)
    and against the host mirror on a batch, including an odd-length curve."""
    g = load_golden("detector")
    det = bf.OddsRatioDetector(_lc(g["a_cts"], g["a_clc"]), noiseestmethod=meth)
    got = det.noise_sigma_batch(g["a_clc"])[0]
    ref = float(g["ref_a_sk_ps" if meth == "powerspectrum" else "ref_a_sk_tv"])
    assert abs(got - ref) <= 1e-12 * ref, (got, ref)
    rng = np.random.default_rng(31)
    for N in (1639, 4400):
        ts = np.arange(N) * DT
        clc = rng.standard_normal((5, N)) * rng.uniform(0.3, 40, (5, 1)) + 3 * np.sin(ts / 2e5)[None, :]
        clc[2] += orc.flare_model(ts, ts[N // 3], 300., 1500., 5000.)
        d = bf.OddsRatioDetector(_lc(ts, clc[0]), noiseestmethod=meth, tvsigma=1.3, psestfrac=0.4)
        dev = d.noise_sigma_batch(clc)
        host = np.array([d.noise_sigma(_lc(ts, clc[b])) for b in range(5)])
        np.testing.assert_allclose(dev, host, rtol=1e-12, atol=0)


def test_batch_with_device_noise_estimate_vs_oracle():
    rng = np.random.default_rng(33)
    B, N = 4, 1200
    ts = np.arange(N) * DT
    clc = rng.standard_normal((B, N)) * np.array([[0.7], [1.0], [5.0], [1.7]])
    clc[1] += orc.flare_model(ts, ts[500], 11., 900., 3000.)
    for meth in ("powerspectrum", "tailveto"):
        det = bf.OddsRatioDetector(_lc(ts, clc[0]), noiseestmethod=meth)
        lnO, sk = det.oddsratio_batch(clc, return_sk=True)
        for b in range(B):
            ref, _, _, sk_ref = orc.oddsratio(clc[b], np.zeros(N), ts, noiseestmethod=meth, return_parts=True)
            assert abs(sk[b] - sk_ref) <= 1e-12 * sk_ref
            assert_logodds_close(lnO[b], ref, TOL, "device noise %s %d" % (meth, b))


# ---- size-independent properties at BASELINE's full sizes ------------------------------------
def test_full_size_batch_properties_config3():
    """1,024 x 4,400 (config 3 on one GPU): (i) every sampled curve of the batch equals its
    single-curve run to rounding, (ii) permuting curves permutes results, (iii) the oracle on two
    curves, (iv) strong injected flares are recovered at their injection time."""
    from bayesflare_b200.synthetic import simulate_batch
    B, N = 1024, 4400
    cts, clc, truth = simulate_batch(B, N, seed=7)
    det = bf.OddsRatioDetector(_lc(cts, clc[0]))
    lnO, sk = det.oddsratio_batch(clc, return_sk=True)
    assert lnO.shape == (B, N - 54) and np.all(np.isfinite(lnO))
    for b in (0, 137, 512, 1023):
        single = det.plan().detector_odds(clc[b], None, sk[b], k0=27, k1=N - 27)
        np.testing.assert_allclose(lnO[b], single, rtol=0, atol=2e-11)   # linear-domain table (batch) vs log-domain table (single)
    # permuting the curves of a batch permutes the output rows (no cross-talk between curves)
    perm = np.random.default_rng(0).permutation(64)
    again = det.plan().detector_odds(clc[perm], None, sk[perm], k0=27, k1=N - 27)
    np.testing.assert_allclose(again, lnO[perm], rtol=0, atol=2e-11)
    for b in (3, 700):
        ref, _ = orc.oddsratio(clc[b], np.zeros(N), cts, sk=sk[b])
        assert_logodds_close(lnO[b], ref, TOL, "config3 curve %d" % b)
        np.testing.assert_array_equal(np.nonzero(lnO[b] > 16.5)[0], np.nonzero(ref > 16.5)[0])
    # injected flares with SNR > 20 are found at their injection time (+-2 samples)
    strong = np.nonzero(truth["snr"] > 20)[0][:64]
    hits = [np.any(lnO[b, truth["idx"][b] - 27 - 2:truth["idx"][b] - 27 + 3] > 16.5) for b in strong]
    assert np.mean(hits) > 0.95


def test_full_size_short_cadence_properties_config2():
    """130,000-sample series (config 2): a time shift of the data shifts the log odds (translation
    invariance of the interior path), and 8 halo'd chunks reproduce the whole series."""
    torch = pytest.importorskip("torch")
    from bayesflare_b200.parallel import time_chunks
    rng = np.random.default_rng(2)
    N = 130000
    ts = np.arange(N) * 58.84876
    clc = rng.standard_normal(N)
    for idx in (20000, 64000, 99000):
        clc[idx - 100:idx + 600] += 9.0 * orc.flare_model(ts[idx - 100:idx + 600], ts[idx], 1., 250., 900.)
    det = bf.OddsRatioDetector(_lc(ts, clc))
    plan, sk = det.plan(), 1.0
    whole = plan.detector_odds(clc, None, sk, k0=27, k1=N - 27)
    assert np.all(np.isfinite(whole))
    shift = 1234
    rolled = plan.detector_odds(np.roll(clc, shift), None, sk, k0=27, k1=N - 27)
    np.testing.assert_allclose(rolled[shift + 60:-60], whole[60:-shift - 60], rtol=0, atol=1e-12)
    dev = torch.device("cuda:0")
    d_sk = torch.tensor([sk], dtype=torch.float64, device=dev)
    out = np.empty(N - 54)
    for k0, k1, s0, s1 in time_chunks(N, 8, 55, True):
        d_seg = torch.tensor(clc[s0:s1], dtype=torch.float64, device=dev)
        d_out = torch.empty(k1 - k0, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        plan.detector_odds_dev(d_seg.data_ptr(), 0, True, d_sk.data_ptr(), 1, N, s0, s1 - s0, k0, k1, d_out.data_ptr())
        plan.ctx.sync()
        out[k0 - 27:k1 - 27] = d_out.cpu().numpy()
    # chunks and the whole series may take different kernel formulations: equal to rounding
    np.testing.assert_allclose(out, whole, rtol=1e-13, atol=2e-12)
    fl, n, mx = det.thresholder(whole, 16.5, expand=1)
    assert n >= 3 and all(any(abs(m[1] + 27 - idx) <= 3 for m in mx) for idx in (20000, 64000, 99000))


def test_injection_sweep_driver_config5_small():
    """BASELINE config 5 mechanics on a small sweep: counters are consistent, strong flares are found,
    and sharding the realisations differently (as 1 or 3 ranks would) gives the same counters."""
    from bayesflare_b200.sweep import detection_efficiency
    whole = detection_efficiency(600, N=1639, seed=4, batch=256)
    assert whole["injected"].sum() == 600 and np.all(whole["detected"] <= whole["injected"])
    assert whole["detected"][25:].sum() >= 0.97 * whole["injected"][25:].sum()
    assert whole["detected"][:3].sum() <= 0.2 * max(whole["injected"][:3].sum(), 1)
    parts = [detection_efficiency(600, N=1639, seed=4, batch=256, rank=r, world=3) for r in range(3)]
    np.testing.assert_array_equal(sum(p["injected"] for p in parts), whole["injected"])
    np.testing.assert_array_equal(sum(p["detected"] for p in parts), whole["detected"])
    assert sum(p["false_alarms"] for p in parts) == whole["false_alarms"]


# ---- per-sample variances (real light curves carry PDCSAP_FLUX_ERR): interior samples take the weighted
# ---- orthonormal-basis kernels (k_window_wsep / k_window_weighted), the edges the reference-order path
def _varying_cle_case(N, seed):
    rng = np.random.default_rng(seed)
    ts = np.arange(N) * DT
    clc = 1.1 * rng.standard_normal(N) + orc.flare_model(ts, ts[N // 2], 8., 1200., 3900.) \
        + 1.5 * np.sin(2 * np.pi * ts / (3.1 * 86400.))
    clc -= np.median(clc)
    cle = 0.35 + 0.25 * rng.random(N) + 0.1 * np.sin(np.arange(N) / 17.)
    return ts, clc, cle


@pytest.mark.gpu
@pytest.mark.parametrize("bglen,bgorder", [(55, 4), (55, 2), (55, 5), (55, 0), (35, 3), (33, 3), (21, 1)])
def test_weighted_detector_vs_oracle(bglen, bgorder):
    """default detector on a curve with per-sample errors, edges included; bglen 21 / 33 / 55 take the split form
    (k_wcorr + k_wsolve), bglen 35 has no separable form (the plain weighted kernel runs), bgorder 0..5 instantiate
    every polynomial size"""
    ts, clc, cle = _varying_cle_case(900, 100 + bgorder)
    kw = dict(bglen=bglen, bgorder=bgorder, noiseestmethod="tailveto", ignoreedges=False)
    if bglen != 55:
        kw["noiseimpulseparams"] = {"t0": (np.inf,)}
    det = bf.OddsRatioDetector(_lc(ts, clc, cle), **kw)
    lnO, _ = det.oddsratio()
    ref, _ = orc.oddsratio(clc, cle, ts, **kw)
    assert_logodds_close(lnO, ref, TOL, "weighted W%d o%d" % (bglen, bgorder))
    np.testing.assert_array_equal(np.nonzero(np.array(lnO) > 10.0)[0], np.nonzero(ref > 10.0)[0])


@pytest.mark.gpu
def test_weighted_dense_grid_multi_pass_and_scripts_detector_vs_oracle():
    """a 24 x 20 flare grid needs several passes over the rise shapes; the scripts' detector adds a
    55-point impulse grid (one-tap supports) and reversed decays"""
    ts, clc, cle = _varying_cle_case(400, 7)
    fp = {"taugauss": (0, 5400, 24), "tauexp": (1800, 10800, 20)}
    kw = dict(flareparams=fp, noiseestmethod="tailveto", noiseimpulseparams={"t0": (0, 54 * DT, 55)})
    det = bf.OddsRatioDetector(_lc(ts, clc, cle), **kw)
    lnO, _ = det.oddsratio()
    ref, _ = orc.oddsratio(clc, cle, ts, **kw)
    assert_logodds_close(lnO, ref, TOL, "weighted dense grid")


@pytest.mark.gpu
def test_weighted_batch_rows_equal_single_curves_and_full_lnB():
    B, N = 6, 700
    curves = [_varying_cle_case(N, 200 + b) for b in range(B)]
    ts = curves[0][0]
    clc = np.stack([c[1] for c in curves])
    cle = np.stack([c[2] for c in curves])
    det = bf.OddsRatioDetector(_lc(ts, clc[0], cle[0]))
    sk = np.array([det.noise_sigma(_lc(ts, clc[b], cle[b])) for b in range(B)])
    batch = det.oddsratio_batch(clc, sk=sk, cle=cle)
    assert batch.shape == (B, N - 54)
    for b in (0, 3, 5):
        ref, _ = orc.oddsratio(clc[b], cle[b], ts, sk=float(sk[b]))
        assert_logodds_close(batch[b], ref, TOL, "weighted batch row %d" % b)
    # the per-grid-point output (Bayes.lnBmargAmp) takes the plain weighted kernel in full mode
    pr = {"t0": (np.inf,), "amp": (1.,), "taugauss": (0, 5400, 5), "tauexp": (1800, 10800, 5)}
    Bf = bf.Bayes(_lc(ts, clc[1], cle[1]), bf.Flare(ts, amp=1, paramranges=pr))
    Bf.bayes_factors_marg_poly_bgd()
    s1 = det.noise_sigma(_lc(ts, clc[1], cle[1]))
    assert det.noiseestmethod == "powerspectrum"
    ref, _, refbg = orc.bayes_factors(clc[1], cle[1], ts, "flare", pr, s1, with_background=True)
    assert_logodds_close(Bf.lnBmargAmp, ref, TOL, "weighted full lnB")
    assert_logodds_close(Bf.bayes_factors_marg_poly_bgd_only(), refbg, TOL, "weighted bg only")


_SPLIT_VS_ONE_KERNEL = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
import bayesflare_b200 as bf
rng = np.random.default_rng(31)
B, N, DT = 5, 700, 1765.55929
ts = np.arange(N) * DT
clc = rng.standard_normal((B, N)) + 2.0 * np.sin(2 * np.pi * ts / (2.7 * 86400.))[None, :]
clc[2, 300:306] += np.array([9., 7., 5., 3.5, 2., 1.])
clc[4, 500] = np.nan
cle = 0.3 + 0.25 * rng.random((B, N))
det = bf.OddsRatioDetector(bf.Lightcurve(cts=ts, clc=clc[0], cle=cle[0]))
np.save(sys.argv[1], det.oddsratio_batch(clc, sk=np.full(B, 0.9), cle=cle))
"""


@pytest.mark.gpu
def test_weighted_split_form_in_chunks_equals_the_one_kernel_form(tmp_path):
    """k_wcorr + k_wsolve with a 1 MB cap on the window-sum scratch (one curve per chunk, five chunks; a flare, so
    the log-domain fallback runs; a NaN) against k_window_wsep -- two processes, the switches are read once"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "run.py"
    script.write_text(_SPLIT_VS_ONE_KERNEL % {"root": root})
    outs = []
    for name, env in (("split", {"BFL_WSPLIT_CAP_MB": "1"}), ("one", {"BFL_NO_WSPLIT": "1"})):
        out = tmp_path / (name + ".npy")
        subprocess.run([sys.executable, str(script), str(out)], check=True, env=dict(os.environ, **env), timeout=300)
        outs.append(np.load(out))
    split, one = outs
    np.testing.assert_array_equal(np.isfinite(split), np.isfinite(one))
    assert np.count_nonzero(~np.isfinite(split[4])) == 55 and np.all(np.isfinite(split[:4]))
    ok = np.isfinite(one)
    assert np.max(np.abs(split[ok] - one[ok]) / np.maximum(1.0, np.abs(one[ok]))) <= 1e-10
    assert one[2].max() > 16.5 and int(np.argmax(split[2])) == int(np.argmax(one[2]))


@pytest.mark.gpu
def test_weighted_path_non_finite_samples():
    """a NaN in the data poisons exactly the windows that contain it (log_marg_amp_full.c:20,25)"""
    ts, clc, cle = _varying_cle_case(600, 9)
    clc = clc.copy()
    clc[300] = np.nan
    det = bf.OddsRatioDetector(_lc(ts, clc, cle))
    got = det.plan().detector_odds(clc, cle, 1.0, k0=27, k1=600 - 27)
    bad = ~np.isfinite(got)
    expect = np.zeros(600 - 54, dtype=bool)
    expect[300 - 27 - 27:300 - 27 + 28] = True
    np.testing.assert_array_equal(bad, expect)


@pytest.mark.gpu
def test_kepler_search_from_fits_files_vs_reference(tmp_path):
    """SURVEY 8(f) f4 end to end: FITS files -> Lightcurve conditioning -> batched device search -> the
    script's JSON record, against the reference's own detector output on the same file content."""
    from bayesflare_b200 import fitsio
    g = load_golden("ingest")
    paths = []
    for i, kid in enumerate((1001, 1002, 1003)):
        p = str(tmp_path / ("kplr%09d_llc.fits" % kid))
        flux = g["flux32"].copy()
        if i == 2:
            flux = np.float32(12000.) + (flux - np.float32(12000.)) * np.float32(0.02)    # a quiet star
        fitsio.write_lightcurve(p, g["time_days"], flux, g["err32"], kepid=kid, quarter=1)
        paths.append(p)
    lc = bf.Lightcurve(curve=paths[0], maxgap=2)
    det = bf.OddsRatioDetector(lc, noiseestmethod='tailveto', tvsigma=1.0)
    det.set_noise_impulse(noiseimpulseparams={'t0': (0, (det.bglen - 1.) * lc.dt(), det.bglen)})
    lnO, ts = det.oddsratio()
    assert_logodds_close(lnO, g["ref_lnO"], TOL, "ingested curve")       # identical NaN pattern around the gaps
    out = bf.flare_search(paths, threshold=16.5, expand=1, star_info={1001: {"teff": 4000, "logg": 4.5}})
    assert out["No. stars analysed"] == 3 and out["Star list"] == [1001, 1002, 1003]
    assert out["Flaring stars"] == [1001, 1002] and out["No. flares"] == 2 * len(g["ref_segs"])
    assert out["Running window length"] == 55 and out["Noise estimation method"] == "tailveto"
    rec = out["Flaring star data"]["KPLR000001001"]
    assert rec["Max. indices"] == [int(i) for i in g["ref_max"][:, 1]] and rec["Teff"] == 4000
    np.testing.assert_allclose(rec["Max. odds ratios"], g["ref_max"][:, 0], rtol=1e-9)
    assert [s[0] for s in rec["Segment indices"]] == [int(a) for a in g["ref_segs"][:, 0]]
    assert [s[-1] + 1 for s in rec["Segment indices"]] == [int(b) for b in g["ref_segs"][:, 1]]
    assert abs(rec["Noise sigma"] - float(g["ref_sk"])) <= 1e-9 * float(g["ref_sk"])
    assert rec["Lightcurve DC background"] == float(g["ref_dc"])
    np.testing.assert_allclose(rec["Max. times"], [g["ref_cts"][27:-27][int(i)] for i in g["ref_max"][:, 1]])
    bf.write_result(out, str(tmp_path / "res.json"))
    # restart: already analysed stars are skipped
    again = bf.flare_search(paths, threshold=16.5, outdict=out)
    assert again["No. stars analysed"] == 3
    # checkpoint file: written after every batch, resumed from when present
    ck = str(tmp_path / "ck.json")
    first = bf.flare_search(paths[:2], threshold=16.5, checkpoint=ck, batch=1)
    assert first["No. stars analysed"] == 2 and os.path.exists(ck)
    resumed = bf.flare_search(paths, threshold=16.5, checkpoint=ck)
    assert resumed["Star list"] == [1001, 1002, 1003] and resumed["No. flares"] == out["No. flares"]
    assert resumed["Flaring star data"].keys() == out["Flaring star data"].keys()


@pytest.mark.gpu
def test_count_regions_with_batch_totals():
    import torch
    rng = np.random.default_rng(3)
    B, n = 7, 500
    lnO = rng.standard_normal((B, n)) * 3
    lnO[2] = -5.0                                    # a curve without detections
    ctx = _lib.default_context()
    d = torch.from_numpy(lnO).cuda()
    d_cnt = torch.empty(B, dtype=torch.int64, device="cuda")
    d_tot = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    for _ in range(2):                               # totals accumulate over calls
        ctx.count_regions_dev(d.data_ptr(), B, n, 2.0, d_cnt.data_ptr(), d_tot.data_ptr())
    ctx.sync()
    want = np.array([len(orc.contiguous_regions(row > 2.0)) for row in lnO])
    np.testing.assert_array_equal(d_cnt.cpu().numpy(), want)
    assert d_tot.cpu().tolist() == [2 * int(want.sum()), 2 * int((want > 0).sum())]


# ---- edge cases: minimum / ragged sizes, widest window, fully masked grids, bad arguments
@pytest.mark.gpu
@pytest.mark.parametrize("N", [55, 56, 60, 109, 110])
def test_shortest_series_vs_oracle(N, monkeypatch):
    """N == bglen (every window truncated) up to N == 2*bglen (first full windows): the reference's
    edge handling bayes.py:320-327 with as few as 28 samples in a window"""
    rng = np.random.default_rng(N)
    ts = np.arange(N) * DT
    clc = rng.standard_normal(N) + orc.flare_model(ts, ts[N // 2], 6., 1000., 3000.)
    _force_sk(monkeypatch, 1.07)
    pr = {"t0": (np.inf,), "amp": (1.,), "taugauss": (0, 5400, 3), "tauexp": (1800, 10800, 3)}
    B = bf.Bayes(_lc(ts, clc), bf.Flare(ts, amp=1, paramranges=pr))
    B.bayes_factors_marg_poly_bgd()
    ref, _, refbg = orc.bayes_factors(clc, np.zeros(N), ts, "flare", pr, 1.07, with_background=True)
    assert B.lnBmargAmp.shape == (1, 3, 3, 1, N)
    assert_logodds_close(B.lnBmargAmp, ref, TOL, "N=%d" % N)
    det = bf.OddsRatioDetector(_lc(ts, clc), ignoreedges=False)
    monkeypatch.setattr(det, "noise_sigma", lambda lightcurve=None: 1.07)
    lnO, tso = det.oddsratio()
    refO, _ = orc.oddsratio(clc, np.zeros(N), ts, ignoreedges=False, sk=1.07)
    assert len(lnO) == N
    assert_logodds_close(lnO, refO, TOL, "detector N=%d" % N)


@pytest.mark.gpu
def test_too_short_series_and_bad_arguments():
    ts = np.arange(54) * DT
    det = bf.OddsRatioDetector(_lc(np.arange(300) * DT, np.zeros(300)))
    plan = det.plan()
    with pytest.raises(_lib.BflError):                      # bayes.py: bglen greater than the data length
        plan.detector_odds(np.zeros(54), None, 1.0)
    with pytest.raises(ValueError):                         # bad sample range
        plan.detector_odds(np.zeros(300), None, 1.0, k0=200, k1=100)
    with pytest.raises(ValueError):                         # ragged batch: sigma count differs from curve count
        plan.detector_odds(np.zeros((3, 300)), None, np.ones(2))
    with pytest.raises(ValueError):
        plan.detector_odds(np.zeros((3, 300)), np.zeros((2, 300)), np.ones(3))
    with pytest.raises(ValueError):                         # neither sigma nor a noise specification
        plan.detector_odds(np.zeros(300), None, None)
    # a sub-range of samples equals the same slice of the full run
    rng = np.random.default_rng(1)
    clc = rng.standard_normal(300)
    full = plan.detector_odds(clc, None, 1.0)
    part = plan.detector_odds(clc, None, 1.0, k0=13, k1=200)
    np.testing.assert_allclose(part, full[13:200], rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
def test_widest_window_and_fully_masked_grid(monkeypatch):
    """bglen = 255 (the ABI's limit) against the oracle; a flare grid whose every point has tau_gauss >
    tau_exp (prior -inf everywhere) marginalises to -inf, like the reference's masked rows"""
    rng = np.random.default_rng(255)
    N = 700
    ts = np.arange(N) * 60.0
    clc = rng.standard_normal(N) + orc.flare_model(ts, ts[350], 5., 600., 2400.)
    _force_sk(monkeypatch, 0.98)
    pr = {"t0": (np.inf,), "amp": (1.,), "taugauss": (0, 1800, 2), "tauexp": (600, 3600, 3)}
    B = bf.Bayes(_lc(ts, clc), bf.Flare(ts, amp=1, paramranges=pr))
    B.bayes_factors_marg_poly_bgd(bglen=255, bgorder=3)
    ref, _ = orc.bayes_factors(clc, np.zeros(N), ts, "flare", pr, 0.98, bglen=255, bgorder=3)
    assert_logodds_close(B.lnBmargAmp, ref, TOL, "bglen 255")
    with pytest.raises(_lib.BflError):
        _lib.Plan(_lib.default_context(), 257, bbayes.background_models(257, 3), np.arange(257) * 60., [])
    prm = {"t0": (np.inf,), "amp": (1.,), "taugauss": (7200, 9000, 2), "tauexp": (600, 3600, 2)}
    Bm = bf.Bayes(_lc(ts, clc), bf.Flare(ts, amp=1, paramranges=prm))
    Bm.bayes_factors_marg_poly_bgd()
    assert np.all(Bm.lnBmargAmp == -np.inf)
    refm, rr = orc.bayes_factors(clc, np.zeros(N), ts, "flare", prm, 0.98)
    assert np.all(refm == -np.inf)


@pytest.mark.gpu
def test_sweep_on_device_simulation_scoring_and_statistics():
    """(i) the device bookkeeping kernel equals the host restatement of the script's
    segment logic on the same log odds; (ii) device-simulated curves have the statistics asked for
    (unit white noise, median removed, injected flare at the recorded index with the recorded SNR);
    (iii) the all-device sweep is independent of the sharding and agrees with the host-simulated sweep
    within counting errors."""
    import torch
    from bayesflare_b200.sweep import detection_efficiency, score_host
    from bayesflare_b200.synthetic import simulate_batch
    ctx = _lib.default_context()
    N, h, nb = 1639, 27, 300
    # (i) scoring
    cts, clc, truth = simulate_batch(nb, N, seed=9)
    truth["idx"][:5] = -1 + 0 * truth["idx"][:5]            # a few curves without an injection record
    det = bf.OddsRatioDetector(_lc(cts, clc[0]), noiseestmethod="tailveto")
    lnO = det.oddsratio_batch(clc)
    edges = np.linspace(0., 50., 51)
    inj_trim = np.where(truth["idx"] >= 0, truth["idx"] - h, -1)
    want = score_host(lnO, inj_trim, truth["snr"], 5.0, edges)     # low threshold: many regions and false alarms
    tr = np.stack([np.where(truth["idx"] >= 0, truth["idx"], -1).astype(float), truth["snr"], truth["taugauss"], truth["tauexp"]], axis=1)
    d_lnO, d_tr = torch.from_numpy(lnO).cuda(), torch.from_numpy(tr).cuda()
    d_inj = torch.zeros(50, dtype=torch.int64, device="cuda"); d_det = torch.zeros_like(d_inj)
    d_cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    ctx.sweep_score_dev(d_lnO.data_ptr(), nb, N - 2 * h, h, d_tr.data_ptr(), 5.0, 50, 50., d_inj.data_ptr(), d_det.data_ptr(), d_cnt.data_ptr())
    ctx.sync()
    np.testing.assert_array_equal(d_inj.cpu().numpy(), want[0])
    np.testing.assert_array_equal(d_det.cpu().numpy(), want[1])
    assert d_cnt.cpu().tolist() == [want[2], want[3]] and want[2] > 10
    # (ii) simulation
    B = 512
    d_clc = torch.empty((B, N), dtype=torch.float64, device="cuda"); d_t = torch.empty((B, 4), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.sweep_simulate_dev(123, 1000, B, N, h, DT, 1.0, 0.0, 0., 50., True, d_clc.data_ptr(), d_t.data_ptr())
    ctx.sync()
    x, t = d_clc.cpu().numpy(), d_t.cpu().numpy()
    assert np.all(np.isfinite(x)) and np.allclose(np.median(x, axis=1), 0.0, atol=1e-12)
    idx, snr, tg, te = t[:, 0].astype(int), t[:, 1], t[:, 2], t[:, 3]
    assert idx.min() >= h and idx.max() <= N - h - 2 and np.all(tg <= te) and np.all((te >= 1800) & (te <= 10800))
    assert 20 < snr.mean() < 30 and snr.min() >= 0 and snr.max() <= 50
    ts = np.arange(N) * DT
    for b in (0, 17, 300):                                   # remove the recorded injection: unit white noise remains
        m = orc.flare_model(ts, ts[idx[b]], 1.0, tg[b], te[b])
        lo, hi = max(0, idx[b] - 4 * h), min(N, idx[b] + 12 * h)
        m[:lo] = 0; m[hi:] = 0
        resid = x[b] - m * snr[b] / np.sqrt(np.sum(m ** 2))
        assert abs(np.std(resid) - 1.0) < 0.08 and abs(np.mean(resid - np.median(resid))) < 0.1
    noise = x[snr < 1.0]
    if len(noise):
        assert abs(np.std(noise) - 1.0) < 0.05
    ctx.sweep_simulate_dev(123, 1100, 8, N, h, DT, 1.0, 0.0, 0., 50., True, d_clc.data_ptr(), d_t.data_ptr())   # realisation 1100.. again
    ctx.sync()
    np.testing.assert_array_equal(d_clc[:8].cpu().numpy(), x[100:108])      # streams depend on (seed, index) only
    # (iii) all-device sweep
    whole = detection_efficiency(3000, N=N, seed=4, batch=1024, device_sim=True)
    assert whole["injected"].sum() == 3000 and np.all(whole["detected"] <= whole["injected"])
    parts = [detection_efficiency(3000, N=N, seed=4, batch=512, rank=r, world=3, device_sim=True) for r in range(3)]
    np.testing.assert_array_equal(sum(p["injected"] for p in parts), whole["injected"])
    np.testing.assert_array_equal(sum(p["detected"] for p in parts), whole["detected"])
    assert sum(p["false_alarms"] for p in parts) == whole["false_alarms"]
    host = detection_efficiency(3000, N=N, seed=4, batch=1024)
    eff_d = whole["detected"][20:].sum() / whole["injected"][20:].sum()
    eff_h = host["detected"][20:].sum() / host["injected"][20:].sum()
    assert eff_d > 0.97 and abs(eff_d - eff_h) < 0.02
    mid_d = whole["detected"][5:15].sum() / whole["injected"][5:15].sum()
    mid_h = host["detected"][5:15].sum() / host["injected"][5:15].sum()
    assert abs(mid_d - mid_h) < 0.08                         # transition region: binomial scatter of ~600 curves


@pytest.mark.gpu
def test_long_series_time_chunks_driver():
    """`longseries.oddsratio_time_chunks`: the pieces 4 ranks would compute, concatenated, are the whole
    series"""
    from bayesflare_b200.longseries import oddsratio_time_chunks
    rng = np.random.default_rng(3)
    N = 30000
    cts = np.arange(N) * 58.84876
    x = rng.standard_normal(N)
    x[12000:12400] += 9.0 * np.exp(-np.arange(400) * 58.84876 / 900.)
    det = bf.OddsRatioDetector(_lc(cts, x))
    whole, _ = oddsratio_time_chunks(det, x, sk=1.0, rank=0, world=1)
    lnO, _ = det.oddsratio()
    assert len(whole) == N - 54
    pieces = [oddsratio_time_chunks(det, x, sk=1.0, rank=r, world=4, gather=False)[0] for r in range(4)]
    # (different sample counts per launch may pick a different samples-per-thread tiling: rounding-level)
    np.testing.assert_allclose(np.concatenate(pieces), whole, rtol=0, atol=2e-12)
    assert abs(int(np.argmax(whole)) - (12000 - 27)) <= 3 and whole.max() > 16.5, (int(np.argmax(whole)), whole.max())
    sk = float(det.noise_sigma_batch(x)[0])
    auto, _ = oddsratio_time_chunks(det, x, rank=0, world=1)          # sigma estimated on the device
    np.testing.assert_allclose(auto, np.array(lnO), rtol=1e-10, atol=1e-10)
    # what the ranks exchange is region records, not log odds: the records of 4 chunks (one region is cut by a
    # chunk boundary on purpose), joined, are thresholder's segments of the whole series
    from bayesflare_b200 import longseries
    thr = 10.0
    cut = longseries.time_chunks(N, 4, 55, True)[1][0]                # first output sample of chunk 1
    y = x + orc.flare_model(cts, cts[cut], 40., 500., 2500.)          # a broad, strong flare peaking ON the boundary
    det2 = bf.OddsRatioDetector(_lc(cts, y))
    recs = []
    for r in range(4):
        run = longseries._runner(det2, N, r, 4)
        run.run(y, 1.0, thresh=thr)
        recs.append(run.regions())
    joined = longseries.join_regions(np.concatenate(recs))
    whole2, _ = oddsratio_time_chunks(det2, y, sk=1.0, rank=0, world=1)
    fl, n, mx = det2.thresholder(whole2, thr, expand=0)
    assert n == len(joined) and n >= 2 and sum(len(r) for r in recs) > n      # a region was split and re-joined
    np.testing.assert_array_equal(joined[:, :2].astype(int), np.array(fl, dtype=int).reshape(-1, 2))
    np.testing.assert_array_equal(joined[:, 2].astype(int), [m[1] for m in mx])
    np.testing.assert_allclose(joined[:, 3], [m[0] for m in mx], rtol=0, atol=2e-12)
    one, _, _ = longseries.regions_time_chunks(det2, y, thr, sk=1.0)
    np.testing.assert_array_equal(one[:, :3], joined[:, :3])


# ---------------------------------------------------------------------------------------------------------
# round 2: parity in the reference's own ill-conditioned regime, transit odds at the sweep's shape, the
# pipelined batch-search ABI
def _one_ulp(x, seed=0):
    rng = np.random.default_rng(seed)
    return np.nextafter(x, np.where(rng.random(x.shape) < 0.5, -np.inf, np.inf))


@pytest.mark.parametrize("case", ["a", "b"])
def test_large_polynomial_component_vs_reference(case):
    """Fixtures tests/golden/trend.npz are outputs of the REFERENCE ITSELF (NumPy + OpenBLAS np.correlate, Cython,
    C) on (a) a slow trend 400 sigma high, (b) sigma = 50 on a DC offset of 3000 -- data whose polynomial-like
    component is far above the noise, the common regime of real PDCSAP light curves.  The reference eliminates raw
    monomials there, so its own answer moves by `floor` when ONE BIT of the input changes (measured here with the
    oracle) and by the same amount between summation orders of the BLAS dot behind np.correlate (the oracle port,
    ascending order, differs from the reference fixture by 4.7e-10 on case a -- tests/test_oracle_golden.py).
    Agreement below a few times that floor is not defined by the reference; the tolerance is therefore
    max(1e-9, 4 floor) and the floor is printed.  Detections above threshold must still be identical."""
    g = load_golden("trend")
    cts, clc, sk = g[case + "_cts"], g[case + "_clc"], float(g[case + "_sk"])
    ref = g["ref_" + case + "_lnO"]
    N = len(clc)
    det = bf.OddsRatioDetector(_lc(cts, clc))
    got = det.plan().detector_odds(clc, None, sk, k0=27, k1=N - 27)
    o1, _ = orc.oddsratio(clc, np.zeros(N), cts, sk=sk)
    o2, _ = orc.oddsratio(_one_ulp(clc), np.zeros(N), cts, sk=sk)
    floor = float(np.max(np.abs(o2 - o1) / np.maximum(1.0, np.abs(o1))))
    err = float(np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))))
    print("case %s: gpu vs reference %.3e, oracle vs reference %.3e, reference's one-ulp floor %.3e"
          % (case, err, float(np.max(np.abs(o1 - ref) / np.maximum(1.0, np.abs(ref)))), floor))
    assert err <= max(TOL, 4.0 * floor), (err, floor)
    for thr in (16.5, 10.0):
        np.testing.assert_array_equal(np.nonzero(got > thr)[0], np.nonzero(ref > thr)[0])


def test_transit_odds_config5_shape_vs_reference(monkeypatch):
    """BASELINE config 5's transit-model odds at the sweep's shape (N = 1639, scripts' settings, tail-veto noise):
    Bayes(lc, Transit).bayes_factors_marg_poly_bgd against the reference's own output (tests/golden/transit5.npz)."""
    g = load_golden("transit5")
    cts, clc = g["cts"], g["clc"]
    lc = _lc(cts, clc)
    assert abs(bf.OddsRatioDetector(lc, noiseestmethod="tailveto").noise_sigma() - float(g["ref_sk"])) < 1e-12
    sg, tf = g["sigmag"], g["tauf"]
    ranges = {"t0": (np.inf,), "sigmag": (sg[0], sg[1], int(sg[2])), "tauf": (tf[0], tf[1], int(tf[2])), "amp": (1.,)}
    B = bf.Bayes(lc, bf.Transit(cts, amp=1, paramranges=ranges))
    B.bayes_factors_marg_poly_bgd(bglen=55, bgorder=4, noiseestmethod="tailveto", tvsigma=1.0, halfrange=True)
    assert B.lnBmargAmp.shape == g["ref_lnB"].shape
    assert_logodds_close(B.lnBmargAmp, g["ref_lnB"], TOL, "transit lnB")
    assert_logodds_close(B.marginalise_full().lnBmargAmp, g["ref_marg"], TOL, "transit marginal")


def test_pe_grid_2d_marginals_vs_reference():
    g = load_golden("pegrid")
    lc = _lc(g["cts"], g["clc"])
    t0, idx, a = float(g["t0"]), int(g["idxt0"]), g["amprange"]
    pe = bf.ParameterEstimationGrid('flare', lc)
    pe.lightcurve_chunk(idx, 55)
    pe.set_grid(ranges={'taugauss': (0., 3600., 12), 'tauexp': (0., 12000., 10), 'amp': (a[0], a[1], int(a[2])), 't0': (t0,)})
    pe.calculate_posterior(bgorder=3)
    for pair in (["taugauss", "tauexp"], ["amp", "taugauss"]):
        got = pe.marginalised_posterior_2D(pair)
        ref = g["ref_marg2d_%s_%s" % tuple(pair)]
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-300)
    assert pe.marginalised_posterior_2D(["taugauss"]) is None and pe.marginalised_posterior("nope") is None


def test_pipelined_batch_search_matches_blocking_call_and_thresholder():
    """bfl_pipe (asynchronous H2D / kernels / D2H, several jobs in flight): log odds identical to the blocking
    bfl_detector_odds, region records identical to thresholder (find.py:915-1019 without expansion) per curve."""
    torch = pytest.importorskip("torch")
    from bayesflare_b200.synthetic import simulate_batch
    B, N = 96, 1639
    cts, clc, _ = simulate_batch(B, N, seed=21)
    det = bf.OddsRatioDetector(_lc(cts, clc[0]))
    plan = det.plan()
    K0, K1 = 27, N - 27
    depth, nj = 3, 7
    # every job gets its own input (rows rotated); the blocking call on the same input is the reference
    # (same kernels, same inputs: bit-identical.  A curve's result may depend on its ROW at the 1e-13 level --
    # the batched FFT of the noise estimate -- so rotated outputs of one blocking call would not do.)
    refs = [det.oddsratio_batch(np.roll(clc, j, axis=0), return_sk=True) for j in range(nj)]
    pipe = _lib.Pipe(plan, B, N, K0, K1, noise=det.noise_spec(), depth=depth, max_regions=8 * B)
    h_in = [torch.from_numpy(np.roll(clc, j, axis=0).copy()).pin_memory() for j in range(nj)]
    h_out = [torch.empty((B, K1 - K0), dtype=torch.float64).pin_memory() for _ in range(depth)]
    h_cnt = [torch.zeros(B + 3, dtype=torch.int64).pin_memory() for _ in range(depth)]
    h_reg = [torch.zeros(8 * B * 3, dtype=torch.float64).pin_memory() for _ in range(depth)]
    h_sk = [torch.zeros(B, dtype=torch.float64).pin_memory() for _ in range(depth)]
    thr = 12.0
    jobs = []
    for j in range(nj):
        s = j % depth
        if j >= depth:
            check_job(pipe, jobs[j - depth], j - depth, depth, h_out, h_cnt, h_reg, h_sk, refs[j - depth], thr, B, det)
        jobs.append(pipe.submit(h_in[j].data_ptr(), None, h_out[s].data_ptr(), thr, h_reg[s].data_ptr(), h_cnt[s].data_ptr(),
                                h_sk[s].data_ptr()))
    for j in range(nj - depth, nj):
        check_job(pipe, jobs[j], j, depth, h_out, h_cnt, h_reg, h_sk, refs[j], thr, B, det)
    assert pipe.launches > 0
    with pytest.raises(_lib.BflError):
        pipe.submit(h_in[0].data_ptr(), None, None, float("nan"), None, None)     # nothing to return


def check_job(pipe, job, j, depth, h_out, h_cnt, h_reg, h_sk, ref, thr, B, det):
    nreg = pipe.wait(job)
    s = j % depth
    lnO = h_out[s].numpy()
    np.testing.assert_array_equal(lnO, ref[0])                    # same kernels, same inputs: bit-identical
    np.testing.assert_array_equal(h_sk[s].numpy(), ref[1])
    cnt = h_cnt[s].numpy()
    rec = np.frombuffer(h_reg[s].numpy().tobytes(), dtype=_lib.REGION_DTYPE)[:nreg]
    rec = np.sort(rec, order=["curve", "start"])
    assert nreg == cnt[B] == cnt[B + 2] and cnt[B + 1] == np.count_nonzero(cnt[:B])
    k = 0
    for b in range(B):
        fl, n, mx = det.thresholder(lnO[b], thr, expand=0)
        assert n == cnt[b]
        for seg, m in zip(fl, mx):
            r = rec[k]
            assert (r["curve"], r["start"], r["stop"], r["argmax"]) == (b, seg[0], seg[1], m[1]) and r["maxval"] == m[0]
            k += 1
    assert k == nreg


def test_kepler_search_with_per_star_time_bases_and_a_bad_file(tmp_path):
    """Real TIME columns are barycentre-corrected per target: every star has its own low bits.  Stars whose
    cadence agrees to cadence_rtol share one device plan (one batch); cadence_rtol=0 gives each its own plan and
    must reproduce the per-star detector exactly; a file that cannot be read is rejected, not fatal."""
    from bayesflare_b200 import fitsio
    g = load_golden("ingest")
    rng = np.random.default_rng(5)
    paths = []
    for i, kid in enumerate((2001, 2002, 2003, 2004)):
        p = str(tmp_path / ("kplr%09d_llc.fits" % kid))
        t = g["time_days"] * (1.0 + 2e-6 * i) + 1e-7 * rng.standard_normal(len(g["time_days"])).cumsum() * 0  # smooth stretch
        fitsio.write_lightcurve(p, t, g["flux32"], g["err32"], kepid=kid, quarter=1)
        paths.append(p)
    bad = str(tmp_path / "kplr000002999_llc.fits")
    with open(bad, "wb") as f:
        f.write(b"not a fits file")
    batched = bf.flare_search(paths + [bad], threshold=16.5, expand=1)
    assert batched["No. stars analysed"] == 4 and len(batched["Rejected stars"]) == 1
    exact = bf.flare_search(paths, threshold=16.5, expand=1, cadence_rtol=0)
    assert exact["No. stars analysed"] == 4 and exact["Flaring stars"] == batched["Flaring stars"] == [2001, 2002, 2003, 2004]
    for kid, p in zip((2001, 2004), (paths[0], paths[3])):
        lc = bf.Lightcurve(curve=p, maxgap=2)
        det = bf.OddsRatioDetector(lc, noiseestmethod='tailveto', tvsigma=1.0)
        det.set_noise_impulse(noiseimpulseparams={'t0': (0, (det.bglen - 1.) * lc.dt(), det.bglen)})
        lnO, _ = det.oddsratio()
        fl, n, mx = det.thresholder(lnO, 16.5, expand=1)
        rec = exact["Flaring star data"]["KPLR%09d" % kid]
        assert rec["Max. indices"] == [m[1] for m in mx]
        np.testing.assert_allclose(rec["Max. odds ratios"], [m[0] for m in mx], rtol=1e-9)
        # the shared-plan batch: same detections, maxima equal to ~cadence_rtol-level shifts of the time base
        recb = batched["Flaring star data"]["KPLR%09d" % kid]
        assert recb["Max. indices"] == rec["Max. indices"]
        np.testing.assert_allclose(recb["Max. odds ratios"], rec["Max. odds ratios"], rtol=1e-4)


def test_sweep_script_recipe_injection_on_device_vs_host():
    """SURVEY 8(f) row f2, the script's recipe to the letter (scripts/flare_detection_efficiency.py:291-335, 429-434,
    444-453, 483-487): flare AND transit injections rescaled with the noise sigma estimated from the base curve.
    (i) The device-injected curve equals the host restatement (models.Flare / models.Transit on the whole series,
    noise.estimate_noise_tv of the Savitzky-Golay residual) applied to the SAME base curve and drawn parameters;
    (ii) the drawn transit parameters obey the script's constraints; (iii) the all-device sweep with transit
    injections is independent of the sharding and agrees with the host-simulated sweep within counting errors."""
    import torch
    from bayesflare_b200.sweep import detection_efficiency
    from bayesflare_b200.synthetic import script_injection_scale
    ctx = _lib.default_context()
    N, h, B = 1639, 27, 64
    ts = np.arange(N) * DT
    spec = _lib.make_noise_spec("tailveto", 55, 4, tvsigma=1.0)
    for kind, mode in ((1, 2), (2, 3)):
        d_clc = torch.empty((B, N), dtype=torch.float64, device="cuda")
        d_t = torch.empty((B, 4), dtype=torch.float64, device="cuda")
        d_sk = torch.empty(B, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        ctx.sweep_simulate_dev(77, 500, B, N, h, DT, 1.0, 2.0, 0., 50., mode, d_clc.data_ptr(), d_t.data_ptr())
        ctx.sync()
        base, t = d_clc.cpu().numpy().copy(), d_t.cpu().numpy()
        ctx.noise_sigma_dev(spec, d_clc.data_ptr(), B, N, d_sk.data_ptr())
        ctx.sweep_inject_dev(kind, B, N, DT, d_sk.data_ptr(), d_t.data_ptr(), d_clc.data_ptr())
        ctx.sync()
        got, sk_dev = d_clc.cpu().numpy(), d_sk.cpu().numpy()
        idx, snr, p1, p2 = t[:, 0].astype(int), t[:, 1], t[:, 2], t[:, 3]
        assert idx.min() >= h and idx.max() <= N - h - 2 and snr.min() >= 0 and snr.max() <= 50
        if kind == 2:
            assert np.all(p2 + 4 * p1 <= 43200.0) and np.all((p2 >= 1800) & (p2 <= 14400)) and np.all((p1 >= 0) & (p1 <= 10800))
        else:
            assert np.all(p1 <= p2) and np.all((p2 >= 1800) & (p2 <= 10800))
        for b in (0, 5, 33, 63):
            if kind == 1:
                inj = bf.Flare(ts, amp=1.).model({"t0": ts[idx[b]], "amp": 1., "taugauss": p1[b], "tauexp": p2[b]})
            else:
                inj = bf.Transit(ts, amp=1.).model({"t0": ts[idx[b]], "amp": 1., "sigmag": p1[b], "tauf": p2[b]})
            fac, sk_host = script_injection_scale(base[b], inj, snr[b])
            assert abs(sk_dev[b] - sk_host) <= 1e-12 * sk_host
            want = base[b] + inj * fac
            np.testing.assert_allclose(got[b], want, rtol=0, atol=1e-11 * max(1.0, np.max(np.abs(want))))
    # --injtransit (script :616-623): every above-threshold segment of the flare detector is a false positive
    whole = detection_efficiency(2048, N=N, seed=6, batch=1024, device_sim=True, recipe="script", inject="transit", sinusoid_amp=2.0)
    assert whole["injected"].sum() == 0 and whole["detected"].sum() == 0 and whole["false_alarms"] > 0
    parts = [detection_efficiency(2048, N=N, seed=6, batch=512, rank=r, world=2, device_sim=True, recipe="script",
                                  inject="transit", sinusoid_amp=2.0) for r in range(2)]
    assert sum(p["false_alarms"] for p in parts) == whole["false_alarms"]
    # device and host simulations of the same recipe must agree statistically (different random streams)
    host = detection_efficiency(512, N=N, seed=6, batch=512, recipe="script", inject="transit", sinusoid_amp=2.0)
    fd, fh = whole["false_alarms"] / 2048.0, host["false_alarms"] / 512.0
    assert abs(fd - fh) < 0.15 * max(fd, fh), (fd, fh)
    fl_d = detection_efficiency(1024, N=N, seed=8, batch=1024, device_sim=True, recipe="script", inject="flare")
    fl_h = detection_efficiency(512, N=N, seed=8, batch=512, recipe="script", inject="flare")
    ed = fl_d["detected"][20:].sum() / fl_d["injected"][20:].sum()
    eh = fl_h["detected"][20:].sum() / fl_h["injected"][20:].sum()
    assert ed > 0.97 and abs(ed - eh) < 0.03


WHOLE_MODELS = {
    "flare": (bf.Flare, {"t0": (np.inf,), "taugauss": (0, 5400, 3), "tauexp": (1800, 10800, 3), "amp": (1.,)}, True),
    "expdecay": (bf.Expdecay, {"t0": (np.inf,), "tauexp": (0., 900., 3), "amp": (1.,)}, True),
    "step": (bf.Step, {"t0": (np.inf,), "amp": (1.,)}, False),
}


@pytest.mark.parametrize("tag", ["odd", "even"])
def test_whole_series_background_vs_reference(tag, monkeypatch):
    """SURVEY row f3, the executable half: bglen=None, nsinusoids=0 (bayes.py:292-294, 330-331, 394-395, 408-409) --
    one polynomial background over the whole series, the template slid along it -- against the reference's own
    output (tests/golden/whole.npz; odd and even series lengths: the reference's end slices place the template one
    sample late for an even length, and some samples are NaN where the truncated template is empty or zero).
    Also the detector built on it (find.py:741-864 with bglen=None: no edges dropped)."""
    g = load_golden("whole")
    cts, clc, cle, sk, bgorder = g[tag + "_cts"], g[tag + "_clc"], g[tag + "_cle"], float(g[tag + "_sk"]), int(g[tag + "_bgorder"])
    _force_sk(monkeypatch, sk)
    lc = _lc(cts, clc, cle)
    with pytest.raises(ValueError):            # the reference raises too unless the curve is already detrended
        bf.Bayes(lc, bf.Step(cts, amp=1, paramranges=dict(WHOLE_MODELS["step"][1]))).bayes_factors_marg_poly_bgd(bglen=None)
    lc.detrended = True
    marg = {}
    for name, (cls, pr, hr) in WHOLE_MODELS.items():
        B = bf.Bayes(lc, cls(cts, amp=1, paramranges=dict(pr)))
        B.bayes_factors_marg_poly_bgd(bglen=None, bgorder=bgorder, halfrange=hr)
        assert_logodds_close(B.lnBmargAmp, g["ref_%s_lnB_%s" % (tag, name)], TOL, "whole-series " + name)
        marg[name] = B.marginalise_full().lnBmargAmp
        assert_logodds_close(marg[name], g["ref_%s_marg_%s" % (tag, name)], TOL, "whole-series marginal " + name)
    bgo = B.bayes_factors_marg_poly_bgd_only(bglen=None, bgorder=bgorder)
    assert_logodds_close(bgo, g["ref_%s_bgonly" % tag], TOL, "whole-series polynomial only")
    with pytest.raises(NotImplementedError):
        B.bayes_factors_marg_poly_bgd(bglen=None, nsinusoids=2)
    det = bf.OddsRatioDetector(lc, bglen=None, bgorder=bgorder, flareparams=dict(WHOLE_MODELS["flare"][1]),
                               noiseimpulse=False, noiseexpdecay=True, noiseexpdecayparams={"tauexp": (0., 900., 3)},
                               noiseexpdecaywithreverse=False, noisestep=True)
    lnO, ts = det.oddsratio()
    assert len(lnO) == len(cts) and np.array_equal(ts, cts)
    want = []
    for i in range(len(cts)):
        den = -np.inf
        for arr in (g["ref_%s_bgonly" % tag], g["ref_%s_marg_expdecay" % tag], g["ref_%s_marg_step" % tag]):
            den = orc.logplus(den, arr[i])
        want.append(g["ref_%s_marg_flare" % tag][i] - den)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(lnO), fin)
    err = np.abs(np.array(lnO)[fin] - np.array(want)[fin]) / np.maximum(1.0, np.abs(np.array(want)[fin]))
    assert err.max() <= TOL, err.max()


@pytest.mark.parametrize("bglen,bgorder", [(21, 2), (33, 4), (101, 4)])
def test_split_form_other_window_lengths_vs_oracle(bglen, bgorder):
    """The two-kernel split form (k_corr_shapes<W, KT> + k_epilogue2) is instantiated for W = 21, 33, 55, 101: a batch
    large enough to take it, against the oracle on two curves and against the single-kernel path curve by curve."""
    from bayesflare_b200.synthetic import simulate_batch
    B, N = 192, 1500
    cts, clc, _ = simulate_batch(B, N, seed=31)
    det = bf.OddsRatioDetector(_lc(cts, clc[0]), bglen=bglen, bgorder=bgorder)
    h = bglen // 2
    sk = np.full(B, 1.07)
    lnO = det.oddsratio_batch(clc, sk=sk)
    assert lnO.shape == (B, N - 2 * h) and np.all(np.isfinite(lnO))
    for b in (0, B - 1):
        ref, _ = orc.oddsratio(clc[b], np.zeros(N), cts, sk=1.07, bglen=bglen, bgorder=bgorder)
        assert_logodds_close(lnO[b], ref, TOL, "split form W=%d curve %d" % (bglen, b))
        single = det.plan().detector_odds(clc[b], None, 1.07, k0=h, k1=N - h)
        np.testing.assert_allclose(lnO[b], single, rtol=0, atol=5e-11)
