"""Kepler flare search driver -- the loop of ``scripts/kepler_analysis_script.py:340-455`` with the
per-star work batched on the device, producing the script's JSON result format.

The script runs, per light-curve file: ``Lightcurve(curve=kfile, maxgap=2)``, the detector
``OddsRatioDetector(lc, noiseestmethod='tailveto', tvsigma=1.0)`` with the impulse hypothesis on every
window sample (``:376-377``), ``oddsratio``, ``thresholder(lnO, threshold, expand=1)`` (``:379-382``), and
for stars with detections a record of maxima, segments, noise sigma and DC level (``:399-445``).  Here
curves that share a time base (same length and cadence) are stacked and go through
``OddsRatioDetector.oddsratio_batch`` -- noise estimate, every hypothesis, marginalisation and the odds
combination in one device pass, per-sample errors included -- and only thresholding output returns to
the host.  Independent stars shard across ranks with :func:`bayesflare_b200.parallel.shard_range`; the
result dictionaries merge with :func:`merge_results`.
"""
from __future__ import annotations

import datetime
import json

import numpy as np

from .finder import OddsRatioDetector
from .lightcurve import Lightcurve

__all__ = ["flare_search", "merge_results", "new_result", "write_result"]


def new_result(threshold, nstars=0, **meta):
    """The script's output dictionary (kepler_analysis_script.py:294-315), selection-cut fields taken
    from ``meta`` when given."""
    out = {
        "Analysis": "Kepler flare search",
        "Analyser": meta.get("analyser", " "),
        "Notes": meta.get("notes", " "),
        "Condition flag requirement": "None",
        "Teff requirement": meta.get("kic_teff"),
        "log(g) requirement": meta.get("kic_logg"),
        "Cadence requirement": meta.get("ktc_target_type"),
        "Kepler quarter": meta.get("sci_data_quarter"),
        "Total no. of stars": meta.get("ntot", nstars),
        "No. stars post-condition flag veto": meta.get("npreperiodveto", nstars),
        "No. stars post-period veto": meta.get("npostperiodveto", nstars),
        "No. stars analysed": 0,
        "Period limit": meta.get("periodlim"),
        "log odds ratio threshold": threshold,
        "Star list": [],
        "Flaring stars": [],
        "Flaring star data": {},
        "No. flares": 0,
        "Rejected stars": [],
    }
    out["Date"] = datetime.datetime.now(datetime.timezone.utc).replace(tzinfo=None).isoformat(' ')
    return out


def _star_record(kepid, info, lnO, ts, flarelist, nflares, maxlist, sk, dc):
    """kepler_analysis_script.py:399-441"""
    lnO = np.asarray(lnO)
    segidxs = [np.arange(a, b).tolist() for a, b in flarelist]
    return {
        "Kepler ID": kepid,
        "No. flares": nflares,
        "Teff": info.get("teff"),
        "log(g)": info.get("logg"),
        "Max. odds ratios": [m[0] for m in maxlist],
        "Max. indices": [m[1] for m in maxlist],
        "Max. times": [float(ts[m[1]]) for m in maxlist],
        "Segment odds ratios": [lnO[i].tolist() for i in segidxs],
        "Segment indices": segidxs,
        "Segment times": [np.asarray(ts)[i].tolist() for i in segidxs],
        "Noise sigma": float(sk),
        "Lightcurve DC background": float(dc),
    }


def flare_search(curves, threshold=16.5, expand=1, maxgap=2, batch=256, star_info=None, outdict=None,
                 detector_kwargs=None, device=None, checkpoint=None, cadence_rtol=1e-4):
    """Search light curves for flares.

    Stars are batched by time base: same number of samples and the same cadence to within ``cadence_rtol``
    (relative).  Real Kepler TIME columns are barycentre-corrected per target, so two stars of one quarter never
    share their time stamps exactly; the templates of a batch are evaluated on the window times of the batch's
    FIRST star, which moves a star's log odds by ~``cadence_rtol`` x (window length / time scale) relative to the
    reference run on that star alone.  ``cadence_rtol=0`` gives every distinct time base its own plan (the
    reference's arithmetic exactly, one device plan per star).  A file that cannot be read or conditioned goes
    to ``"Rejected stars"`` instead of stopping the search.

    ``curves``: file names of Kepler light-curve FITS files and/or :class:`Lightcurve` objects.
    ``star_info``: optional ``{kepid: {"teff": .., "logg": ..}}``.  ``outdict``: a result dictionary to
    continue (stars already in its ``"Star list"`` are skipped, as the script does on restart).
    ``checkpoint``: a file name; the dictionary is written there after every batch (the script rewrites its
    output file as it goes, ``:343-351``) and, if the file exists and ``outdict`` is not given, the search
    resumes from it.  Returns the result dictionary (see :func:`new_result`)."""
    if checkpoint is not None and outdict is None:
        try:
            with open(checkpoint) as f:
                outdict = json.load(f)
        except (OSError, ValueError):
            outdict = None
    star_info = star_info or {}
    out = outdict if outdict is not None else new_result(threshold, nstars=len(curves))
    done = set(out["Star list"])
    kw = dict(noiseestmethod='tailveto', tvsigma=1.0)
    kw.update(detector_kwargs or {})
    if device is not None:
        kw["device"] = device

    groups = {}                       # (N, cadence bucket) -> light curves that share a time base
    for c in curves:
        try:
            lc = c if isinstance(c, Lightcurve) else Lightcurve(curve=c, maxgap=maxgap)
            bad = lc.datagap or len(lc.clc) < 3 or not np.all(np.isfinite(lc.cts[:3])) or not (lc.dt() > 0)
        except Exception as exc:      # unreadable / malformed file: the script would have stopped here
            name = getattr(c, "id", c)
            print("Rejected %s: %s" % (name, exc))
            out["Rejected stars"].append(name if isinstance(name, (int, str)) else str(name))
            continue
        if bad:
            out["Rejected stars"].append(lc.id)
            continue
        if lc.id in done:
            continue
        done.add(lc.id)
        dt = float(lc.dt())
        key = (len(lc.clc), dt) if not cadence_rtol else (len(lc.clc), int(round(np.log(dt) / cadence_rtol)))
        groups.setdefault(key, []).append(lc)

    first = "Noise estimation method" not in out
    for (N, _), lcs in groups.items():
        det = OddsRatioDetector(lcs[0], **kw)
        if "noiseimpulseparams" not in (detector_kwargs or {}):   # kepler_analysis_script.py:377
            det.set_noise_impulse(noiseimpulseparams={'t0': (0, (det.bglen - 1.) * lcs[0].dt(), det.bglen)})
        if first:                     # kepler_analysis_script.py:384-391
            out["Noise estimation method"] = det.noiseestmethod
            out["Running window length"] = det.bglen
            out["Polynomial background order"] = det.bgorder
            out["Flare taug"] = list(det.flareparams['taugauss'])
            out["Flare taue"] = list(det.flareparams['tauexp'])
            out["Amplitude priors"] = list(det.flareparams['amp'])
            first = False
        h = int(det.bglen / 2) if det.ignoreedges else 0
        for b0 in range(0, len(lcs), batch):
            part = lcs[b0:b0 + batch]
            # a sub-series' templates are evaluated on the first curve's time stamps; curves of one group
            # share the sampling, so the (relative) window times agree
            clc = np.stack([lc.clc for lc in part])
            cle = np.stack([lc.cle for lc in part])
            anyerr = bool(np.any(cle != 0.0))
            if det._device_noise_ok():
                lnO, sks = det.oddsratio_batch(clc, cle=cle if anyerr else None, return_sk=True)
            else:
                sks = np.array([det.noise_sigma(lc) for lc in part])
                lnO = det.oddsratio_batch(clc, sk=sks, cle=cle if anyerr else None)
            for i, lc in enumerate(part):
                out["No. stars analysed"] += 1
                out["Star list"].append(lc.id)
                ts = lc.cts[h:N - h] if h else lc.cts
                flarelist, nfl, maxlist = det.thresholder(lnO[i], threshold, expand=expand)
                if nfl > 0:
                    out["No. flares"] += nfl
                    out["Flaring stars"].append(lc.id)
                    out["Flaring star data"]["KPLR%09d" % int(lc.id)] = _star_record(
                        lc.id, star_info.get(lc.id, {}), lnO[i], ts, flarelist, nfl, maxlist, sks[i], lc.dc)
            if checkpoint is not None:
                write_result(out, checkpoint)
    return out


def merge_results(results):
    """Merge the per-rank dictionaries of a sharded search into one (rank order)."""
    results = [r for r in results if r]
    out = dict(results[0])
    for key in ("Star list", "Flaring stars", "Rejected stars"):
        out[key] = [x for r in results for x in r[key]]
    out["Flaring star data"] = {k: v for r in results for k, v in r["Flaring star data"].items()}
    out["No. flares"] = sum(r["No. flares"] for r in results)
    out["No. stars analysed"] = sum(r["No. stars analysed"] for r in results)
    for key in ("Total no. of stars", "No. stars post-condition flag veto", "No. stars post-period veto"):
        if all(isinstance(r.get(key), (int, np.integer)) for r in results):
            out[key] = int(sum(r[key] for r in results))
    return out


def write_result(outdict, path):
    """``json.dump(outdict, f, indent=2)`` (kepler_analysis_script.py:447-452); NumPy scalars become
    plain numbers."""
    def default(o):
        if isinstance(o, np.integer):
            return int(o)
        if isinstance(o, np.floating):
            return float(o)
        if isinstance(o, np.ndarray):
            return o.tolist()
        raise TypeError(type(o))
    with open(path, "w") as f:
        json.dump(outdict, f, indent=2, default=default)


def main(argv=None):
    """``python -m bayesflare_b200.search 'Q1_public/kplr*_llc.fits' --out flares.json``; under ``torchrun`` the
    files are sharded over the ranks (one GPU each) and rank 0 merges and writes the result."""
    import argparse
    import glob
    import os
    from .parallel import gather_objects, shard_range
    ap = argparse.ArgumentParser(description="Kepler flare search (scripts/kepler_analysis_script.py) on the GPU")
    ap.add_argument("files", nargs="+", help="light-curve FITS files or glob patterns")
    ap.add_argument("--threshold", type=float, default=16.5)
    ap.add_argument("--expand", type=int, default=1)
    ap.add_argument("--maxgap", type=int, default=2)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--out", default="flare_search.json")
    ap.add_argument("--checkpoint", default=None, help="per-rank checkpoint file prefix: written after every batch, resumed from")
    args = ap.parse_args(argv)
    files = sorted(f for pat in args.files for f in (glob.glob(pat) or [pat]))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lo, hi = shard_range(len(files), world, rank)
    res = flare_search(files[lo:hi], threshold=args.threshold, expand=args.expand, maxgap=args.maxgap, batch=args.batch,
                       device=local_rank if world > 1 else None,
                       checkpoint=("%s.rank%d.json" % (args.checkpoint, rank)) if args.checkpoint else None)
    allres = gather_objects(res)
    if rank == 0:
        merged = merge_results(allres)
        write_result(merged, args.out)
        print("%d stars analysed on %d GPU(s), %d flares in %d stars -> %s"
              % (merged["No. stars analysed"], world, merged["No. flares"], len(merged["Flaring stars"]), args.out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
