"""Many spectra per call: Hough + vote + RANSAC + refit for a whole batch
(BASELINE configs[3]: 10k synthetic arcs sharded over the GPUs).

Each spectrum is what one `Calibrator` would hold: peaks, atlas lines, pixel list
and the Hough / RANSAC / fit properties.  The stages are the same C-ABI calls the
single-spectrum facade makes, with `n_problems = len(spectra)`:

    K1 rb2_hough_accumulate -> K2 rb2_candidate_vote -> host set-up (calibrator.py:441-523,
    vectorised) -> K3 rb2_ransac_batch -> K4 rb2_refit_winners -> one D2H copy.

Spectra are independent: with torch.distributed, rank r takes spectra r, r+world, ...
(no data-path collective; results are gathered by the caller if it wants them in one place).
"""
from __future__ import annotations

import math
import types

import numpy as np

from . import engine, pipeline

HOUGH_DEFAULTS = dict(num_slopes=2000, xbins=100, ybins=100, min_wavelength=3000.0, max_wavelength=9000.0,
                      range_tolerance=500, linearity_tolerance=100)
RANSAC_DEFAULTS = dict(sample_size=5, top_n_candidate=5, filter_close=False, ransac_tolerance=5,
                       candidate_weighted=True, hough_weight=1.0, minimum_matches=0,
                       minimum_peak_utilisation=0.0, minimum_fit_error=1e-4)
FIT_DEFAULTS = dict(fit_deg=4, candidate_tolerance=2.0)
_EMPTY = {}


def _limits(h, span):
    """calibrator.py:1088-1116; span = np.ptp(pixel_list)."""
    wmin, wmax, rt, lt = h["min_wavelength"], h["max_wavelength"], h["range_tolerance"], h["linearity_tolerance"]
    min_i = wmin - rt
    max_i = wmin + rt
    lo = ((wmax - rt - lt) - (min_i + rt + lt)) / span
    hi = ((wmax + rt + lt) - (min_i - rt - lt)) / span
    return lo, hi, min_i, max_i


def _spline_tables(hist0, gc0, gcn):
    """pp-form of the 1-D interpolating cubic splines of hist[:, 0] for many spectra at once.

    hist0: [n_spec, xb] first-intercept-row counts; gc0 / gcn: first / last gradient-bin centre.
    The FITPACK interpolating spline (knots x0*4, x[2:-2], xn*4) is the not-a-knot spline, and
    spline interpolation commutes with the affine map of the equally spaced centres onto
    0..xb-1, so ONE scipy call handles the whole batch.  Returns breaks[n_spec, n_pp+1] and
    coef[n_spec, n_pp, 4] in the variable (t - break)."""
    from scipy import interpolate

    n_spec, xb = hist0.shape
    b = interpolate.make_interp_spline(np.arange(xb, dtype=np.float64), hist0.T, k=3)
    brk = np.unique(b.t)  # [n_pp+1] on the integer grid: 0, 2, 3, ..., xb-3, xb-1
    left = brk[:-1]
    step = (gcn - gc0) / (xb - 1)  # grid units -> gradient units
    breaks = gc0[:, None] + brk[None, :] * step[:, None]
    breaks[:, -1] = gcn
    coef = np.empty((n_spec, len(left), 4))
    fact = (1.0, 1.0, 2.0, 6.0)
    for k in range(4):  # Taylor coefficients at the left break, u = (t - break) / step
        coef[:, :, k] = b(left, nu=k).T / fact[k] / step[:, None] ** k
    return breaks, coef


def _sampling_cdf(peaks):
    """calibrator.py:521-523 incl. the index -1 wrap-around; returns uint32 CDF."""
    return engine.sampling_cdf32(peaks)[1]


def ransac_setups(specs, hough, vote):
    """calibrator.py:441-523 for every spectrum of a batch (host, vectorised where it pays)."""
    n = len(specs)
    rows = hough.first_rows()
    by_xb = {}
    for i, sp in enumerate(specs):
        by_xb.setdefault(int(sp["hough"]["xbins"]), []).append(i)
    tables = [None] * n
    for xb, idxs in by_xb.items():
        h0 = np.stack([rows[i] for i in idxs])
        xe = np.stack([hough.xedges(i) for i in idxs])
        ye0 = np.array([hough.yedges(i)[:2] for i in idxs])
        hx = (xe[:, 1] - xe[:, 0]) / 2.0
        hy = (ye0[:, 1] - ye0[:, 0]) / 2.0
        gc0, gcn = xe[:, 1] - hx, xe[:, -1] - hx
        ic0 = ye0[:, 1] - hy
        breaks, coef = _spline_tables(h0, gc0, gcn)
        for k, i in enumerate(idxs):
            tables[i] = (gc0[k], gcn[k], ic0[k], h0[k, 0], h0[k, -1], breaks[k], coef[k].reshape(-1))
    out = []
    for i, sp in enumerate(specs):
        r, f = sp["ransac"], sp["fit"]
        cp, ca = vote.candidates(i)
        if r["filter_close"]:
            uy = np.unique(ca)
            close = uy[:-1][(uy[1:] - uy[:-1]) < 3 * r["ransac_tolerance"]]
            keep = ~np.isin(ca, close)
            cp, ca = cp[keep], ca[keep]
        peaks, counts = np.unique(cp, return_counts=True)
        s = types.SimpleNamespace()
        s.peaks = peaks
        s.cand_off = np.zeros(len(peaks) + 1, dtype=np.int32)
        np.cumsum(counts, out=s.cand_off[1:])
        s.cand_wave = ca  # vote output is already grouped by ascending peak, original order inside
        s.uwaves, wid = np.unique(ca, return_inverse=True)
        s.cand_wid = wid.astype(np.int32)
        s.deg, s.sample_size = int(f["fit_deg"]), int(sp["sample_size"])
        s.enough = len(peaks) > s.deg and len(peaks) > s.sample_size
        s.error = None
        try:
            s.cdf32 = _sampling_cdf(peaks) if s.enough else np.zeros(len(peaks), dtype=np.uint32)
        except ValueError as e:  # "probabilities contain NaN": this spectrum fails, the batch goes on
            s.enough, s.error, s.cdf32 = False, e, np.zeros(len(peaks), dtype=np.uint32)
        s.known_x = s.known_y = np.zeros(0)
        s.tol = float(r["ransac_tolerance"])
        s.min_intercept, s.max_intercept = float(sp["limits"][2]), float(sp["limits"][3])
        ptp = peaks[-1] - peaks[0] if len(peaks) else 0.0
        s.pix_min = float(peaks[0] - ptp * 0.2) if len(peaks) else 0.0
        s.pix_max = float(peaks[-1] + ptp * 0.2) if len(peaks) else 0.0
        s.hough_weight = float(r["hough_weight"])
        s.gc_min, s.gc_max, s.ic_min, s.h_lo, s.h_hi, s.pp_breaks, s.pp_coef = tables[i]
        pl = sp["pixel_list"] if sp["pixel_list"] is not None else np.arange(sp["n_pix"], dtype=np.float64)
        s.n_pix = len(pl)
        s.pixel_list = pl
        s.pix_is_arange = sp["pix_is_arange"]
        s.pix_sorted = True
        mm = min(int(r["minimum_matches"]), len(sp["lines"]), len(sp["peaks"]))
        s.min_inliers = max(s.deg + 1, mm)
        s.min_inliers_frac = float(r["minimum_peak_utilisation"]) * len(sp["peaks"])
        s.cost_ceiling = 1e50
        s.minimum_fit_error = float(r["minimum_fit_error"])
        out.append(s)
    return out


def normalise_spec(sp):
    """Fill defaults, derive limits / pairs; mirrors Calibrator.__init__ + the three setters."""
    peaks = np.asarray(sp["peaks"], dtype=np.float64)
    lines = np.asarray(sp["lines"], dtype=np.float64)
    pl = sp.get("pixel_list")
    if pl is not None:
        pl = np.asarray(pl, dtype=np.float64)
        is_ar = bool(len(pl) and pl[0] == 0 and pl[-1] == len(pl) - 1 and np.all(np.diff(pl) == 1))
        n_pix, span = len(pl), (float(np.ptp(pl)) if len(pl) else 0.0)
    else:  # pixel_list = np.arange(num_pix) (calibrator.py:914-937), materialised only where it is read
        npix = sp.get("num_pix")
        n_pix = math.ceil(npix if npix is not None else 1.1 * peaks.max())
        is_ar, span = True, float(n_pix - 1)
    h = {**HOUGH_DEFAULTS, **sp.get("hough", _EMPTY)}
    r = {**RANSAC_DEFAULTS, **sp.get("ransac", _EMPTY)}
    f = {**FIT_DEFAULTS, **sp.get("fit", _EMPTY)}
    s = min(int(r["sample_size"]), len(lines))
    if s <= f["fit_deg"]:
        s = f["fit_deg"] + 1
    return dict(peaks=peaks, lines=lines, pixel_list=pl, n_pix=n_pix, pix_is_arange=is_ar, hough=h, ransac=r, fit=f,
                sample_size=s, limits=_limits(h, span))


class BatchResult(types.SimpleNamespace):
    pass


def _batch_parts(n):
    """Parts a device-pipeline batch of n spectra is cut into (RB2_BATCH_PARTS overrides; 1 below 512)."""
    import os

    if n < 512:
        return 1
    try:
        return max(1, min(16, int(os.environ.get("RB2_BATCH_PARTS", "2"))))
    except ValueError:
        return 2


def peaks_from_spectra(spectra):
    """The step before the path for a batch: entries that carry an arc `spectrum`
    instead of `peaks` get them from K6 - scipy.signal.find_peaks(**entry["find_peaks"]) followed by
    rascal.util.refine_peaks(window_width=entry.get("window_width", 10)), as the reference's examples
    do (examples/example_lt_sprat_xe.py:27-28) - in one device batch per distinct recipe; `num_pix`
    defaults to len(spectrum), `max_peaks` (capacity per spectrum) to 1024.  Entries that already have `peaks` pass through untouched."""
    from . import util

    out = list(spectra)
    groups = {}
    for i, sp in enumerate(out):
        if sp.get("peaks") is None and sp.get("spectrum") is not None:
            fp = sp.get("find_peaks", {})
            key = (fp.get("height"), fp.get("distance"), fp.get("prominence"), int(sp.get("window_width", 10)),
                   sp.get("max_peaks", 1024))  # output capacity per spectrum; exceeding it raises
            groups.setdefault(key, []).append(i)
    for (h, d, p, ww, cap), idx in groups.items():
        found = util.find_and_refine([out[i]["spectrum"] for i in idx], height=h, distance=d, prominence=p,
                                     window_width=ww, max_peaks=cap)
        for i, pk in zip(idx, found):
            out[i] = dict(out[i], peaks=np.sort(pk))  # refined blends can cross; the path wants ascending peaks
            out[i].setdefault("num_pix", len(out[i]["spectrum"]))
    return out


def calibrate_batch(spectra, n_hypotheses=10_000, seed=0, refit_mode=0, timings=None, match_tolerance=None,
                    device_pipeline=True):
    """do_hough_transform() + fit() for every spectrum of `spectra` in batched device calls.

    All spectra of one call must share fit_deg and sample size (one K3 instantiation).
    Returns a list of BatchResult(fit_coeff, matched_peaks, matched_atlas, rms, residual,
    peak_utilisation, atlas_utilisation, valid, cost, counters); fit_coeff is None where the
    reference would raise "Couldn't fit".

    device_pipeline: batches that qualify (pipeline.eligible: shared Hough / RANSAC shape, increasing
    peaks and lines, pixel_list = arange) run device-resident end to end (rascal_b200.pipeline: one
    H2D, one D2H, K3's tables built by rb2_ransac_prepare); others take the staged path below.

    An entry may carry an arc `spectrum` (+ `find_peaks` = dict(height=, distance=, prominence=),
    `window_width`) instead of `peaks`: peaks_from_spectra() finds and refines them on the device first.

    match_tolerance: if set, Calibrator.match_peaks(tolerance=match_tolerance) follows every fit
    (K5), reading the refit coefficients straight from K4's device records; the result gains
    `match` = BatchResult(fit_coeff, matched_peaks, matched_atlas, rms, residuals, n_ambiguous)."""
    import functools
    import time

    t0 = time.perf_counter()
    raw_specs = peaks_from_spectra(spectra)
    parts = _batch_parts(len(raw_specs)) if device_pipeline else 1
    kw = dict(seed=seed, refit_mode=refit_mode, match_tolerance=match_tolerance)
    if parts > 1:
        # large batch: consecutive parts, each normalised, packed and launched without waiting for the
        # device, so the host prepares part k+1 (and unpacks part k-1) while the device works on part k;
        # a part the device pipeline cannot take goes through the staged path when its turn comes
        cut = [len(raw_specs) * k // parts for k in range(parts + 1)]
        tms = [{} for _ in range(parts)]
        pending = []
        for k in range(parts):
            sk = [normalise_spec(sp) for sp in raw_specs[cut[k]:cut[k + 1]]]
            if pipeline.eligible(sk):
                pending.append(pipeline.run_async(sk, n_hypotheses, timings=tms[k], slot=k, **kw))
            else:
                pending.append(functools.partial(_calibrate_staged, sk, n_hypotheses, timings=tms[k], **kw))
        out = pipeline.ChainedResults([finish() for finish in pending])
        if timings is not None:
            timings.update(pack=sum(t.get("pack", 0.0) for t in tms), unpack=sum(t.get("unpack", 0.0) for t in tms),
                           parts=parts, total=time.perf_counter() - t0)
        return out
    specs = [normalise_spec(sp) for sp in raw_specs]
    if device_pipeline and pipeline.eligible(specs):
        return pipeline.run(specs, n_hypotheses, timings=timings, **kw)
    return _calibrate_staged(specs, n_hypotheses, timings=timings, **kw)


def _calibrate_staged(specs, n_hypotheses, seed=0, refit_mode=0, match_tolerance=None, timings=None):
    """calibrate_batch for normalised specs through the staged engine classes (host RANSAC set-up)."""
    import time

    torch = engine._torch()
    t0 = time.perf_counter()
    hp, vs = [], []
    for sp in specs:
        n_p, n_a = len(sp["peaks"]), len(sp["lines"])
        px = np.repeat(sp["peaks"], n_a)
        py = np.tile(sp["lines"], n_p)
        lo, hi, mi, ma = sp["limits"]
        h = sp["hough"]
        hp.append(dict(pair_x=px, pair_y=py, min_slope=lo, max_slope=hi, min_intercept=mi, max_intercept=ma,
                       n_slopes=int(h["num_slopes"]), xbins=int(h["xbins"]), ybins=int(h["ybins"])))
        sigma = (h["range_tolerance"] + h["linearity_tolerance"]) * 1.1775
        vs.append(dict(pairs=None, peaks=sp["peaks"], lines=sp["lines"], top_n=sp["ransac"]["top_n_candidate"],
                       weighted=sp["ransac"]["candidate_weighted"], tol=sp["fit"]["candidate_tolerance"],
                       gauss_denom=2 * sigma ** 2 + 1e-9))
    hough = engine.HoughBatch(hp).run()
    vote = engine.VoteBatch(hough, vs).run()
    t1 = time.perf_counter()
    hough.prefetch()
    vote.prefetch()
    setups = ransac_setups(specs, hough, vote)
    t2 = time.perf_counter()
    live = [i for i, s in enumerate(setups) if s.enough]
    results = [BatchResult(fit_coeff=None, matched_peaks=[], matched_atlas=[], rms=1e50, residual=None,
                           peak_utilisation=0.0, atlas_utilisation=0.0, valid=False, cost=float("inf"),
                           counters=None, match=None, error=s.error) for s in setups]
    if live:
        rb = engine.RansacBatch([setups[i] for i in live])
        rb.run(0, int(n_hypotheses), seed=seed)
        rb.refit_async(mode=refit_mode)
        mb = None
        if match_tolerance is not None:
            stride = engine.C.sizeof(engine.N.RefitResult)
            mb = engine.MatchBatch([dict(peaks=specs[i]["peaks"], lines=specs[i]["lines"],
                                         d_coeffs=rb.d_refit.data_ptr() + k * stride, n_coef=setups[i].deg + 1,
                                         fit_deg=setups[i].deg, tolerance=match_tolerance, robust_refit=True)
                                    for k, i in enumerate(live)]).run(mode=refit_mode)
        win, fit, mx, my, rs = rb.fetch()
        if mb is not None:
            mres, mmx, mmy, mrs = mb.fetch()
        for k, i in enumerate(live):
            o, w, sp = fit[k], win[k], specs[i]
            res = results[i]
            res.counters = engine.counters_dict(w)
            res.cost = w.cost
            if w.hyp_id >= 0 and o.accepted:
                m, d = o.n_inliers, setups[i].deg
                res.fit_coeff = np.array(o.coeffs[:d + 1])
                res.matched_peaks, res.matched_atlas = mx[k, :m].copy(), my[k, :m].copy()
                res.rms, res.residual = o.rms, rs[k, :m].copy()
                res.peak_utilisation = m / len(sp["peaks"])
                res.atlas_utilisation = m / len(sp["lines"])
                res.valid = m > d + 1
                if mb is not None and mres[k].refit.accepted:
                    q, mm = mres[k], mres[k].n_matched
                    res.match = BatchResult(fit_coeff=np.array(q.refit.coeffs[:d + 1]), matched_peaks=mmx[k, :mm].copy(),
                                            matched_atlas=mmy[k, :mm].copy(), rms=q.refit.rms,
                                            residuals=mrs[k, :mm].copy(), n_ambiguous=q.n_ambiguous)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(hough_vote=t1 - t0, setup=t2 - t1, ransac_refit=t3 - t2)
    return results
