"""Light-curve ingest: the NumPy FITS reader and the reference's conditioning of
file data (data.py:205-351) against a fixture produced by the reference's own methods.  CPU only."""
import json
import os

import numpy as np
import pytest

import bayesflare_b200 as bf
from bayesflare_b200 import fitsio
from bayesflare_b200.search import merge_results, new_result, write_result
from conftest import load_golden


def _write(tmp_path, g, name="kplr000757450-2009166043257_llc.fits", **kw):
    path = os.path.join(str(tmp_path), name)
    fitsio.write_lightcurve(path, g["time_days"], g["flux32"], g["err32"], **kw)
    return path


def test_fits_round_trip_and_layout(tmp_path):
    g = load_golden("ingest")
    path = _write(tmp_path, g, kepid=757450, quarter=3)
    assert os.path.getsize(path) % 2880 == 0
    primary, hdr, rec = fitsio.read_header_and_table(path)
    assert primary["KEPLERID"] == 757450 and primary["OBSMODE"] == "long cadence" and primary["QUARTER"] == 3
    assert hdr["XTENSION"] == "BINTABLE" and hdr["NAXIS2"] == len(g["time_days"]) and hdr["TFIELDS"] == 7
    assert rec.dtype.names == ("TIME", "TIMECORR", "CADENCENO", "SAP_FLUX", "PDCSAP_FLUX", "PDCSAP_FLUX_ERR", "SAP_QUALITY")
    np.testing.assert_array_equal(rec["TIME"], g["time_days"])
    np.testing.assert_array_equal(rec["PDCSAP_FLUX"], g["flux32"])          # NaN == NaN in assert_array_equal
    np.testing.assert_array_equal(rec["PDCSAP_FLUX_ERR"], g["err32"])
    np.testing.assert_array_equal(rec["CADENCENO"], np.arange(len(rec)))
    # the bare three-column layout reads too
    path2 = _write(tmp_path, g, name="bare.fits", extra_columns=False, cadence="short")
    d = fitsio.read_kepler_lightcurve(path2)
    assert d["OBSMODE"] == "short cadence"
    np.testing.assert_array_equal(d["PDCSAP_FLUX"], g["flux32"])


def test_fits_tform_repeat_counts_and_strings(tmp_path):
    """columns with repeat counts, strings and 64-bit integers (other TFORM codes of real files)"""
    dt = np.dtype([("A", ">f8", (3,)), ("NAME", "S5"), ("K", ">i8"), ("FLAG", "i1")])
    rec = np.zeros(4, dt)
    rec["A"] = np.arange(12).reshape(4, 3)
    rec["NAME"] = [b"ab", b"cde", b"f", b"ghijk"]
    rec["K"] = [1, -2, 2 ** 40, 7]
    cards = [fitsio._card("XTENSION", "BINTABLE"), fitsio._card("BITPIX", 8), fitsio._card("NAXIS", 2),
             fitsio._card("NAXIS1", dt.itemsize), fitsio._card("NAXIS2", 4), fitsio._card("PCOUNT", 0),
             fitsio._card("GCOUNT", 1), fitsio._card("TFIELDS", 4)]
    for i, (n, f) in enumerate((("A", "3D"), ("NAME", "5A"), ("K", "1K"), ("FLAG", "L")), 1):
        cards += [fitsio._card("TTYPE%d" % i, n), fitsio._card("TFORM%d" % i, f)]
    primary = [fitsio._card("SIMPLE", True), fitsio._card("BITPIX", 8), fitsio._card("NAXIS", 0),
               fitsio._card("OBJECT", "it's"), fitsio._card("EXPOSURE", 1.5)]
    data = rec.tobytes()
    path = os.path.join(str(tmp_path), "t.fits")
    with open(path, "wb") as f:
        f.write(fitsio._header_bytes(primary) + fitsio._header_bytes(cards) + data + b"\0" * (-len(data) % 2880))
    p, h, r = fitsio.read_header_and_table(path)
    assert p["OBJECT"] == "it's" and p["EXPOSURE"] == 1.5
    np.testing.assert_array_equal(r["A"], rec["A"])
    assert list(r["NAME"]) == [b"ab", b"cde", b"f", b"ghijk"] and list(r["K"]) == [1, -2, 2 ** 40, 7]


def test_lightcurve_from_file_matches_reference_conditioning(tmp_path):
    g = load_golden("ingest")
    lc = bf.Lightcurve(curve=_write(tmp_path, g, kepid=1234567, quarter=2))
    assert lc.id == 1234567 and lc.cadence == "long" and lc.quarter == "2"
    assert lc.clc.dtype == np.float64 and lc.datagap == bool(g["ref_datagap"])
    np.testing.assert_array_equal(lc.clc, g["ref_clc"])       # interpolation incl. its float32 rounding, median removal
    np.testing.assert_array_equal(lc.cts, g["ref_cts"])
    np.testing.assert_array_equal(lc.cle, g["ref_cle"])       # errors keep their NaNs
    assert lc.dc == float(g["ref_dc"])
    assert abs(lc.dt() - 1765.55929) < 1e-4
    # a second quarter of the same star appends; another star is refused (data.py:245-248)
    lc.add_data(curve=_write(tmp_path, g, name="q3.fits", kepid=1234567, quarter=3))
    assert len(lc.clc) == 2 * len(g["ref_clc"]) and lc.quarter == "23"
    with pytest.raises(NameError):
        lc.add_data(curve=_write(tmp_path, g, name="other.fits", kepid=42))


def test_ingest_errors(tmp_path):
    with pytest.raises(NameError):
        bf.Lightcurve(curve=os.path.join(str(tmp_path), "missing.fits"))
    bad = os.path.join(str(tmp_path), "bad.fits")
    with open(bad, "wb") as f:
        f.write(b"not a fits file" * 300)
    with pytest.raises(NameError):
        bf.Lightcurve(curve=bad)
    with pytest.raises(NameError):
        bf.Lightcurve().add_data()


def test_gap_checker_is_bug_compatible():
    """data.py:298-303 compares len(boolean array) with len(y): never True"""
    lc = bf.Lightcurve()
    d = np.ones(50)
    d[10:30] = np.nan
    assert lc.gap_checker(d, maxgap=1) is False


def test_result_dictionary_format_and_merge(tmp_path):
    a = new_result(16.5, nstars=3, kic_teff="<=5150")
    for key in ("Analysis", "Star list", "Flaring stars", "Flaring star data", "No. flares", "Rejected stars",
                "log odds ratio threshold", "No. stars analysed", "Date", "Teff requirement"):
        assert key in a
    b = new_result(16.5, nstars=2)
    a["Star list"] += [1, 2, 3]; a["No. stars analysed"] = 3; a["No. flares"] = 2; a["Flaring stars"] = [2]
    a["Flaring star data"]["KPLR000000002"] = {"No. flares": 2, "Noise sigma": np.float64(1.5), "Max. indices": [np.int64(4)]}
    b["Star list"] += [9, 8]; b["No. stars analysed"] = 2
    m = merge_results([a, b])
    assert m["Star list"] == [1, 2, 3, 9, 8] and m["No. stars analysed"] == 5 and m["No. flares"] == 2
    assert m["Total no. of stars"] == 5
    path = os.path.join(str(tmp_path), "out.json")
    write_result(m, path)
    back = json.load(open(path))
    assert back["Flaring star data"]["KPLR000000002"]["Max. indices"] == [4]


def test_detrenders_vs_reference():
    g = load_golden("ingest")
    np.testing.assert_array_equal(bf.running_median(g["ref_clc"], 21), g["ref_running_median_21"])
    lc = bf.Lightcurve(cts=g["ref_cts"], clc=g["ref_clc"])
    lc.detrend(method="runningmedian", nbins=21)
    np.testing.assert_array_equal(lc.clc, g["ref_clc"] - g["ref_running_median_21"])
    assert lc.detrended and lc.detrend_method == "runningmedian"
    hp = bf.Lightcurve(cts=g["ref_cts"], clc=g["ref_clc"])
    hp.detrend(method="highpassfilter", knee=1. / (0.3 * 86400.))
    assert len(hp.clc) == len(g["ref_clc"]) and np.all(np.isfinite(hp.clc))
    assert np.std(hp.clc) < np.std(g["ref_clc"])               # the slow trend is gone
    with pytest.raises(ValueError):
        bf.Lightcurve(cts=g["ref_cts"], clc=g["ref_clc"]).detrend(method="runningmedian")
    with pytest.raises(ValueError):
        bf.Lightcurve(cts=g["ref_cts"], clc=g["ref_clc"]).detrend(method="nonsense", nbins=5)
