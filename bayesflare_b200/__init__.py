"""bayesflare_b200 -- B200-native (sm_100a) sliding-window Bayesian odds-ratio flare detector.

A drop-in for the hot path of BayesFlare (``OddsRatioDetector.oddsratio`` ->
``Bayes.bayes_factors_marg_poly_bgd`` -> ``log_marg_amp_full_*``, the grid marginalisation and
``ParameterEstimationGrid``): same class names and signatures, same log-odds arrays, computed
by hand-written CUDA kernels behind the C ABI in ``include/bfl.h``.  There is no CPU fallback.

    import bayesflare_b200 as bf
    lc = bf.Lightcurve(cts=ts, clc=flux)           # or any object with cts/clc/cle/detrend()
    lnO, t = bf.OddsRatioDetector(lc).oddsratio()
"""
from .lightcurve import Lightcurve
from .models import Model, Flare, Transit, Expdecay, Impulse, Gaussian, Step, ModelCurve
from .noise import (estimate_noise_ps, estimate_noise_tv, savitzky_golay, nextpow2, running_median,
                    highpass_filter_lightcurve)
from .general import logplus, logminus, logtrapz
from .bayes import Bayes, ParameterEstimationGrid
from .finder import OddsRatioDetector, contiguous_regions
from .search import flare_search, merge_results, write_result
from . import fitsio

__version__ = "0.1.0"

__all__ = ["Lightcurve", "Model", "Flare", "Transit", "Expdecay", "Impulse", "Gaussian", "Step", "ModelCurve",
           "estimate_noise_ps", "estimate_noise_tv", "savitzky_golay", "nextpow2", "running_median",
           "highpass_filter_lightcurve", "logplus", "logminus",
           "logtrapz", "Bayes", "ParameterEstimationGrid", "OddsRatioDetector", "contiguous_regions",
           "flare_search", "merge_results", "write_result", "fitsio"]
