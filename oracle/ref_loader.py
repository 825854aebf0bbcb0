"""Import the *unmodified-on-disk* Python reference from ``/root/reference`` (TEST
INFRASTRUCTURE, container only -- ``/root/reference`` does not exist on the GPU box).

Used by ``tests/golden/make_golden.py`` to generate golden vectors and by the CPU tests
(when the reference is present) to validate the oracle restatement.  Never imported by
the product package.

The reference is Python-2 era code (README.md:19-23).  To run it on Python 3.12 /
NumPy 2.3 without touching the read-only tree, the loader

* serves the package's modules from ``/root/reference/bayesflare`` through an import
  hook that applies the in-memory text patches listed in ``PATCHES``:
  ``ts == None`` -> ``ts is None`` (model.py, elementwise-compare semantics changed),
  an ``int()`` around a float slice index (noise.py:39), ``== None`` on ndarrays in the
  PE post-processing (bayes.py), and the two undefined names in ``Transit.prior``
  (model.py:471,489-492,515; without them Transit odds raise NameError);
* maps ``bayesflare.stats.general`` to the compiled ``oracle/_ref/general*.so``
  (built by ``oracle/build_ref.py`` from the reference's own .pyx/.c);
* stubs the absent third-party imports (matplotlib[.mlab/.pyplot/.gridspec], pyfits,
  astropy.io.fits) -- only ``matplotlib.mlab.psd`` is ever called on the hot path
  (data.py:200) and is restated here with modern-mlab semantics;
* restores removed NumPy/SciPy aliases the source uses (``np.product``, ``np.int``,
  ``np.mat``, ``scipy.signal.boxcar``).
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

from . import build_ref

REF_ROOT = build_ref.REF_ROOT
PKG_DIR = os.path.join(REF_ROOT, "bayesflare")
if not os.path.isdir(PKG_DIR) and os.path.isdir(os.path.join(build_ref.SITE, "bayesflare")):
    # the GPU box: /root/reference does not exist there; oracle/build_ref.install_python() put an unmodified
    # copy of the reference's Python package under oracle/_ref/site (git-ignored, shipped with the snapshot)
    REF_ROOT = build_ref.SITE
    PKG_DIR = os.path.join(REF_ROOT, "bayesflare")
PKG = "bayesflare"

# (module, old, new, expected minimum number of replacements)
PATCHES = {
    "bayesflare.models.model": [
        ("ts == None", "ts is None", 8),
        ("2.*sigmagcutoff*pdict", "2.*sigmacutoff*pdict", 1),
        ("2.*sigmagcutoff )", "2.*sigmacutoff )", 4),
        ("taufranges[0]", "taufrange[0]", 1),
    ],
    "bayesflare.noise.noise": [
        ("sk[np.floor((1.-estfrac)*len(sk)):]", "sk[int(np.floor((1.-estfrac)*len(sk))):]", 1),
    ],
    "bayesflare.stats.bayes": [
        ("self.posterior == None", "self.posterior is None", 3),
        ("self.maxposterior == None", "self.maxposterior is None", 2),
        # Python 2 orders None below every int, so `None > N` is simply False there (bglen=None path, row f3)
        ("if bglen > N:", "if bglen is not None and bglen > N:", 1),
    ],
}


def available() -> bool:
    built = build_ref.build() if build_ref.reference_present() else {"general": None}
    so = built.get("general") or os.path.join(build_ref.OUT, "general" + build_ref.ext_suffix())
    return os.path.isdir(PKG_DIR) and os.path.exists(so)


def mlab_psd(x, NFFT=None, Fs=2, window=None, noverlap=0, sides="onesided", **_):
    """Restatement of modern ``matplotlib.mlab.psd`` for the single call the reference
    makes (data.py:200): one segment of length NFFT=len(x), no detrend, given window,
    ``scale_by_freq=True``: ``|rfft(w x)|^2 / (Fs sum(w^2))``, doubled except DC and
    (for even NFFT) Nyquist."""
    x = np.asarray(x, dtype=float)
    n = len(x) if NFFT is None else int(NFFT)
    w = np.ones(n) if window is None else np.asarray(window, dtype=float)
    seg = x[:n] * w
    spec = np.abs(np.fft.rfft(seg, n=n)) ** 2
    spec /= (np.abs(w) ** 2).sum()
    if n % 2:
        spec[1:] *= 2.0
    else:
        spec[1:-1] *= 2.0
    spec /= Fs
    freqs = np.fft.rfftfreq(n, 1.0 / Fs)
    return spec, freqs


def _install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__dict__["__bfl_stub__"] = True
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    try:  # never shadow a real installation
        import matplotlib  # noqa: F401
        import matplotlib.mlab  # noqa: F401
    except Exception:
        mpl = mod("matplotlib")
        mpl.__path__ = []
        mpl.mlab = mod("matplotlib.mlab", psd=mlab_psd)
        mpl.pyplot = mod("matplotlib.pyplot")
        mpl.gridspec = mod("matplotlib.gridspec")
    try:
        import astropy.io.fits  # noqa: F401
    except Exception:
        ap = mod("astropy")
        ap.__path__ = []
        io = mod("astropy.io")
        io.__path__ = []
        io.fits = mod("astropy.io.fits")
        ap.io = io
    if "pyfits" not in sys.modules:
        try:
            import pyfits  # noqa: F401
        except Exception:
            mod("pyfits")

    # aliases removed from NumPy / SciPy since the reference was written
    if not hasattr(np, "product"):
        np.product = np.prod          # bayes.py:123,347,417,977
    if not hasattr(np, "int"):
        np.int = int                  # noise.py:234-235
    if not hasattr(np, "mat"):
        np.mat = np.asmatrix          # noise.py:245
    import scipy.signal
    if not hasattr(scipy.signal, "boxcar"):
        scipy.signal.boxcar = scipy.signal.windows.boxcar   # data.py:200


class _PatchedLoader(importlib.abc.SourceLoader):
    def __init__(self, fullname, path):
        self.fullname, self.path = fullname, path

    def get_filename(self, fullname):
        return self.path

    def get_data(self, path):
        with open(path, "rb") as fh:
            data = fh.read()
        if path != self.path:
            return data
        text = data.decode("utf-8")
        for old, new, nmin in PATCHES.get(self.fullname, []):
            cnt = text.count(old)
            if cnt < nmin:
                raise ImportError("reference patch %r expected >=%d sites in %s, found %d"
                                  % (old, nmin, self.fullname, cnt))
            text = text.replace(old, new)
        return text.encode("utf-8")

    # never write/read .pyc for the read-only tree
    def path_stats(self, path):
        raise OSError

    def set_data(self, path, data):
        pass


class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != PKG and not fullname.startswith(PKG + "."):
            return None
        if fullname == PKG + ".stats.general":
            built = build_ref.build()
            so = built.get("general")
            if not so:
                raise ImportError("oracle/_ref/general*.so is not built")
            loader = importlib.machinery.ExtensionFileLoader(fullname, so)
            return importlib.util.spec_from_file_location(fullname, so, loader=loader)
        rel = fullname.split(".")[1:]
        base = os.path.join(PKG_DIR, *rel)
        if os.path.isdir(base):
            fn = os.path.join(base, "__init__.py")
            return importlib.util.spec_from_file_location(
                fullname, fn, loader=_PatchedLoader(fullname, fn),
                submodule_search_locations=[base])
        fn = base + ".py"
        if os.path.isfile(fn):
            return importlib.util.spec_from_file_location(
                fullname, fn, loader=_PatchedLoader(fullname, fn))
        return None


_loaded = None


def load_reference():
    """Return the imported reference package ``bayesflare`` (patched in memory)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError("reference tree not present at %s" % REF_ROOT)
    _install_stubs()
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    import importlib
    _loaded = importlib.import_module(PKG)
    return _loaded


class RefLightcurve:
    """Minimal stand-in for ``bayesflare.Lightcurve`` *construction* (the reference's
    constructor only reads FITS files, data.py:126-133): builds a real reference
    ``Lightcurve`` instance and fills ``cts/clc/cle`` directly, as the reference's own
    scripts do (scripts/flare_detection_efficiency.py:416-424)."""

    def __new__(cls, cts, clc, cle=None):
        bf = load_reference()
        lc = bf.Lightcurve()
        lc.cts = np.array(cts, dtype=float)
        lc.clc = np.array(clc, dtype=float)
        lc.cle = np.zeros(len(lc.cts)) if cle is None else np.array(cle, dtype=float)
        return lc
