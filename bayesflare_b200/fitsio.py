"""Minimal FITS reader/writer for Kepler light-curve files (NumPy only).

The reference reads light curves with ``pyfits``/``astropy.io.fits`` (``data.py:233-258``):
primary-header keywords ``KEPLERID``, ``OBSMODE``, ``QUARTER`` and the columns ``TIME``,
``PDCSAP_FLUX``, ``PDCSAP_FLUX_ERR`` of the first extension (a ``BINTABLE``).  Neither library is a
dependency here, and the format is small: 2880-byte blocks, 80-character header cards, big-endian
fixed-width records.  :func:`read_table` understands every fixed-width ``TFORM`` code
(``L X B I J K A E D C M`` with repeat counts), which covers the twenty columns of a real
``*_llc.fits`` / ``*_slc.fits`` file; :func:`write_lightcurve` produces files of the same layout for
tests and synthetic data.
"""
from __future__ import annotations

import re

import numpy as np

__all__ = ["read_header_and_table", "read_kepler_lightcurve", "write_lightcurve"]

BLOCK = 2880
_TFORM = {"L": "i1", "X": "u1", "B": "u1", "I": ">i2", "J": ">i4", "K": ">i8", "A": "S1", "E": ">f4", "D": ">f8",
          "C": ">c8", "M": ">c16"}


def _parse_value(text):
    text = text.strip()
    if text.startswith("'"):
        end = 1
        out = []
        while end < len(text):                    # '' inside a string is an escaped quote
            if text[end] == "'":
                if end + 1 < len(text) and text[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(text[end])
            end += 1
        return "".join(out).rstrip()
    text = text.split("/", 1)[0].strip()
    if text in ("T", "F"):
        return text == "T"
    if not text:
        return None
    try:
        return int(text)
    except ValueError:
        try:
            return float(text.replace("D", "E"))
        except ValueError:
            return text


def _read_header(buf, pos):
    """Header unit starting at byte ``pos``: ``(dict, position of the data unit)``."""
    hdr = {}
    while True:
        if pos + BLOCK > len(buf):
            raise IOError("truncated FITS header")
        block = buf[pos:pos + BLOCK]
        pos += BLOCK
        for i in range(0, BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                return hdr, pos
            if card[8:10] == "= ":
                hdr[key] = _parse_value(card[10:])


def _data_bytes(hdr):
    naxis = int(hdr.get("NAXIS", 0))
    if naxis == 0:
        return 0
    n = abs(int(hdr["BITPIX"])) // 8 * int(hdr.get("GCOUNT", 1))
    size = int(hdr.get("PCOUNT", 0))
    prod = 1
    for i in range(1, naxis + 1):
        prod *= int(hdr["NAXIS%d" % i])
    return n * (size + prod)


def _table_dtype(hdr):
    fields = []
    for i in range(1, int(hdr["TFIELDS"]) + 1):
        m = re.match(r"\s*(\d*)([LXBIJKAEDCM])", str(hdr["TFORM%d" % i]))
        if not m:
            raise IOError("unsupported TFORM%d = %r" % (i, hdr["TFORM%d" % i]))
        rep = int(m.group(1)) if m.group(1) else 1
        code = m.group(2)
        name = str(hdr.get("TTYPE%d" % i, "col%d" % i))
        if code == "A":
            fields.append((name, "S%d" % max(rep, 1)))
        elif code == "X":
            fields.append((name, "u1", ((rep + 7) // 8,)))
        elif rep == 1:
            fields.append((name, _TFORM[code]))
        else:
            fields.append((name, _TFORM[code], (rep,)))
    dt = np.dtype(fields)
    if dt.itemsize != int(hdr["NAXIS1"]):
        raise IOError("row width %d does not match NAXIS1 = %d" % (dt.itemsize, int(hdr["NAXIS1"])))
    return dt


def read_header_and_table(path, ext=1):
    """Primary header and the ``ext``-th extension of ``path`` as ``(primary, ext_header, records)``;
    ``records`` is a NumPy structured array (big-endian fields, one per ``TTYPE``)."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:6] != b"SIMPLE":
        raise IOError("%s is not a FITS file" % path)
    primary, pos = _read_header(buf, 0)
    pos += -(-_data_bytes(primary) // BLOCK) * BLOCK
    hdr = None
    for _ in range(ext):
        hdr, pos = _read_header(buf, pos)
        nbytes = _data_bytes(hdr)
        start = pos
        pos += -(-nbytes // BLOCK) * BLOCK
    if hdr is None or str(hdr.get("XTENSION", "")).strip() != "BINTABLE":
        raise IOError("extension %d of %s is not a binary table" % (ext, path))
    nrows = int(hdr["NAXIS2"])
    dt = _table_dtype(hdr)
    if start + nrows * dt.itemsize > len(buf):
        raise IOError("truncated FITS table")
    return primary, hdr, np.frombuffer(buf, dtype=dt, count=nrows, offset=start)


def read_kepler_lightcurve(path):
    """What ``Lightcurve.add_data`` takes from a Kepler file (data.py:240-256)."""
    primary, _, rec = read_header_and_table(path, 1)
    return {"KEPLERID": primary["KEPLERID"], "OBSMODE": primary.get("OBSMODE", ""), "QUARTER": primary.get("QUARTER", ""),
            "TIME": rec["TIME"], "PDCSAP_FLUX": rec["PDCSAP_FLUX"], "PDCSAP_FLUX_ERR": rec["PDCSAP_FLUX_ERR"]}


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        v = "%20d" % value
    elif isinstance(value, (float, np.floating)):
        v = "%20s" % repr(float(value)).upper()
    else:
        v = "'%-8s'" % str(value).replace("'", "''")
    card = "%-8s= %s" % (key, v)
    if comment:
        card += " / " + comment
    return ("%-80s" % card)[:80]


def _header_bytes(cards):
    text = "".join(cards) + "%-80s" % "END"
    text += " " * (-len(text) % BLOCK)
    return text.encode("ascii")


def write_lightcurve(path, time_days, flux, flux_err, kepid=1, quarter=1, cadence="long", extra_columns=True):
    """Write a Kepler-style light-curve file: primary HDU with ``KEPLERID/OBSMODE/QUARTER`` and a
    ``LIGHTCURVE`` binary table.  ``extra_columns`` adds ``TIMECORR``, ``CADENCENO``, ``SAP_FLUX`` and
    ``SAP_QUALITY`` so that the wanted columns are neither first nor contiguous, as in real files."""
    n = len(time_days)
    cols = [("TIME", ">f8", "D", np.asarray(time_days, dtype=">f8"))]
    if extra_columns:
        cols += [("TIMECORR", ">f4", "E", np.zeros(n, ">f4")), ("CADENCENO", ">i4", "J", np.arange(n, dtype=">i4")),
                 ("SAP_FLUX", ">f4", "E", np.asarray(flux, dtype=">f4"))]
    cols += [("PDCSAP_FLUX", ">f4", "E", np.asarray(flux, dtype=">f4")),
             ("PDCSAP_FLUX_ERR", ">f4", "E", np.asarray(flux_err, dtype=">f4"))]
    if extra_columns:
        cols += [("SAP_QUALITY", ">i4", "J", np.zeros(n, ">i4"))]
    rec = np.zeros(n, dtype=np.dtype([(c[0], c[1]) for c in cols]))
    for name, _, _, data in cols:
        rec[name] = data
    primary = [_card("SIMPLE", True), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True),
               _card("KEPLERID", int(kepid), "unique Kepler target identifier"),
               _card("OBSMODE", "%s cadence" % cadence, "observing mode"), _card("QUARTER", int(quarter))]
    ext = [_card("XTENSION", "BINTABLE"), _card("BITPIX", 8), _card("NAXIS", 2), _card("NAXIS1", rec.dtype.itemsize),
           _card("NAXIS2", n), _card("PCOUNT", 0), _card("GCOUNT", 1), _card("TFIELDS", len(cols))]
    for i, (name, _, code, _) in enumerate(cols, 1):
        ext += [_card("TTYPE%d" % i, name), _card("TFORM%d" % i, code)]
    ext.append(_card("EXTNAME", "LIGHTCURVE"))
    data = rec.tobytes()
    data += b"\0" * (-len(data) % BLOCK)
    with open(path, "wb") as f:
        f.write(_header_bytes(primary))
        f.write(_header_bytes(ext))
        f.write(data)
