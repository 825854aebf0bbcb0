"""Build the reference's own native code into ``oracle/_ref/`` (TEST INFRASTRUCTURE).

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU-baseline leg may use the
outputs.  Nothing here is imported by the product package ``bayesflare_b200``.

What is built (sources are compiled *where they lie* under ``/root/reference``; no
reference source is copied into this repository):

* ``oracle/_ref/liblogmarg.so``  -- ``bayesflare/stats/log_marg_amp_full.c`` alone, so
  ``log_marg_amp_full_C`` / ``log_marg_amp_except_final_C`` can be called through ctypes.
* ``oracle/_ref/general.<abi>.so`` -- the Cython module ``bayesflare/stats/general.pyx``
  plus the C file, the same Extension ``setup.py:39-51`` declares (minus ``-lgsl``:
  GSL is absent here, ``oracle/shim/gsl`` supplies ``gsl_sf_log_erfc``/``gsl_sf_erf``).

Both are compiled with the reference's own flags (``-O3``, no ``-march``, hence no FMA
contraction on x86-64) so rounding behaviour is the reference's.

``install_python()`` additionally installs the reference's *Python* package, unmodified, into
``oracle/_ref/site/bayesflare`` (what ``pip install --target`` would put there; git-ignored like the
rest of ``oracle/_ref`` but shipped to the GPU box), so that ``bench.py``'s CPU arms can time the
reference's own NumPy path on the box's host cores.  ``oracle/ref_loader.py`` imports it from there
(with its in-memory Python-3 patches) when ``/root/reference`` itself is absent.

The recipe is a no-op when ``/root/reference`` is missing (the GPU box): prebuilt files
travel with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("BFL_REFERENCE_ROOT", "/root/reference")
REF_STATS = os.path.join(REF_ROOT, "bayesflare", "stats")
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "shim")


def reference_present() -> bool:
    return os.path.isfile(os.path.join(REF_STATS, "log_marg_amp_full.c"))


def _run(cmd):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), res.stdout, res.stderr))


def _newer(target, *sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def ext_suffix() -> str:
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


def build(force: bool = False, verbose: bool = False) -> dict:
    """Compile the reference's native sources.  Returns paths of what exists."""
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "liblogmarg.so")
    ext = os.path.join(OUT, "general" + ext_suffix())
    if not reference_present():
        return {"liblogmarg": lib if os.path.exists(lib) else None,
                "general": ext if os.path.exists(ext) else None}

    csrc = os.path.join(REF_STATS, "log_marg_amp_full.c")
    pyx = os.path.join(REF_STATS, "general.pyx")
    shim_h = os.path.join(SHIM, "gsl", "gsl_sf_erf.h")
    cflags = ["-O3", "-fPIC", "-std=gnu99", "-I" + SHIM, "-I" + REF_STATS]

    if force or not _newer(lib, csrc, shim_h):
        _run(["gcc", *cflags, "-shared", csrc, "-o", lib, "-lm"])
        if verbose:
            print("built", lib)

    if force or not _newer(ext, csrc, pyx, shim_h):
        import numpy as np
        bdir = os.path.join(OUT, "build")
        os.makedirs(bdir, exist_ok=True)
        gen_c = os.path.join(bdir, "general.c")
        # language_level=2 as the 2016-era source expects (general.c:1 was Cython 0.24)
        _run([sys.executable, "-m", "cython", "-2", "-I", REF_STATS, "-o", gen_c, pyx])
        inc = ["-I" + sysconfig.get_paths()["include"], "-I" + np.get_include()]
        _run(["gcc", *cflags, "-w", "-shared", *inc,
              "-DNPY_NO_DEPRECATED_API=0", gen_c, csrc, "-o", ext, "-lm"])
        if verbose:
            print("built", ext)
    return {"liblogmarg": lib, "general": ext}


SITE = os.path.join(OUT, "site")


def install_python(force: bool = False) -> str | None:
    """Install the reference's Python package (``*.py`` only, byte for byte) under ``oracle/_ref/site``."""
    import shutil
    src = os.path.join(REF_ROOT, "bayesflare")
    dst = os.path.join(SITE, "bayesflare")
    if not reference_present():
        return dst if os.path.isdir(dst) else None
    if os.path.isdir(dst) and not force:
        return dst
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for root, dirs, files in os.walk(src):
        rel = os.path.relpath(root, src)
        os.makedirs(os.path.join(dst, rel), exist_ok=True)
        for f in files:
            if f.endswith(".py"):
                shutil.copy2(os.path.join(root, f), os.path.join(dst, rel, f))
    return dst


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(install_python(force="--force" in sys.argv))
