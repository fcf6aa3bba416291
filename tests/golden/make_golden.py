#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (it needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

The reference (pure Python, rascal v0.3.9) is imported from
/root/reference/src with two shims that do not change its arithmetic
(SURVEY.md App. B): a stub `pynverse` module, and `do_hough_transform`
re-stated to skip the `self.pairs == []` comparison that raises under
numpy >= 1.25 (it makes the same three HoughTransform calls, calibrator.py:
1434-1452).  Recording hooks wrap `np.polynomial.polynomial.polyfit`,
`Calibrator._match_bijective`, `models.robust_polyfit`, the Hough-density
spline and the logger; nothing else is touched.

Each fixture holds the inputs, the reference's Hough histogram / edges /
point count / point checksum, its own (unstable) bin order, the candidates it
derives under the canonical tie order, and the full seeded RANSAC trace
(every drawn sample as index sets, status, cost, acceptance trail, result).
"""

import hashlib
import json
import logging
import os
import sys
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.modules.setdefault("pynverse", types.ModuleType("pynverse"))
sys.path.insert(0, "/root/reference/src")

import rascal.calibrator as rc  # noqa: E402
from rascal import models  # noqa: E402
from rascal.atlas import Atlas  # noqa: E402
from rascal.calibrator import Calibrator  # noqa: E402
from rascal.houghtransform import HoughTransform  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

ST_ABORT, ST_INTERCEPT, ST_MONOTONIC, ST_EMPTY, ST_SCORED = range(5)


def _do_hough(self, brute_force=False):
    self.ht.set_constraints(self.min_slope, self.max_slope,
                            self.min_intercept, self.max_intercept)
    if brute_force:
        self.ht.generate_hough_points_brute_force(self.pairs[:, 0],
                                                  self.pairs[:, 1])
    else:
        self.ht.generate_hough_points(self.pairs[:, 0], self.pairs[:, 1],
                                      num_slopes=self.num_slopes)
    self.ht.bin_hough_points(self.xbins, self.ybins)
    self.hough_points = self.ht.hough_points
    self.hough_lines = self.ht.hough_lines


Calibrator.do_hough_transform = _do_hough

_orig_bin = HoughTransform.bin_hough_points
_CANON = {"on": False}


def _bin_canonical(self, xbins, ybins):
    """Reference binning; optionally re-rank ties canonically afterwards."""
    _orig_bin(self, xbins, ybins)
    self.ref_order = np.ravel_multi_index(
        (self.hist_sorted_arg[:, 0], self.hist_sorted_arg[:, 1]),
        self.hist.shape)
    if _CANON["on"]:
        order = np.argsort(self.hist.ravel(), kind="stable")[::-1]
        i, j = np.unravel_index(order, self.hist.shape)
        self.hist_sorted_arg = np.column_stack((i, j))
        hx = (self.xedges[1] - self.xedges[0]) / 2
        hy = (self.yedges[1] - self.yedges[0]) / 2
        self.hough_lines = [(self.xedges[a] + hx, self.yedges[b] + hy)
                            for a, b in zip(i, j)]


HoughTransform.bin_hough_points = _bin_canonical


class Recorder(logging.Handler):
    """Collects the event stream of one Calibrator.fit() call."""

    def __init__(self):
        super().__init__(level=logging.DEBUG)
        self.events = []

    def emit(self, record):
        self.events.append(("log", record.getMessage()))

    def install(self, c):
        self.c = c
        c.logger.handlers = [self]
        c.logger.setLevel(logging.DEBUG)
        rec = self
        self._polyfit = np.polynomial.polynomial.polyfit

        def polyfit(x, y, deg, *a, **k):
            out = rec._polyfit(x, y, deg, *a, **k)
            rec.events.append(("fit", np.array(x, float), np.array(y, float),
                               np.array(out)))
            return out

        np.polynomial.polynomial.polyfit = polyfit
        # fit_type 'legendre' / 'chebyshev' (calibrator.py:1625-1631) fit with these instead
        self._legfit, self._chebfit = np.polynomial.legendre.legfit, np.polynomial.chebyshev.chebfit

        def wrap(fn):
            def fit(x, y, deg, *a, **k):
                out = fn(x, y, deg, *a, **k)
                rec.events.append(("fit", np.array(x, float), np.array(y, float), np.array(out)))
                return out
            return fit

        np.polynomial.legendre.legfit = wrap(self._legfit)
        np.polynomial.chebyshev.chebfit = wrap(self._chebfit)
        self._mb = Calibrator._match_bijective

        def mb(self_, cands, peaks, coeff):
            out = rec._mb(self_, cands, peaks, coeff)
            rec.cand_dict = cands
            rec.ransac_peaks = np.array(peaks)
            rec.events.append(("match", out[0].copy(), out[1].copy(),
                               out[2].copy()))
            return out

        Calibrator._match_bijective = mb
        self._rp = models.robust_polyfit

        def rp(x, y, degree=3, x0=None):
            out = rec._rp(x, y, degree, x0)
            rec.events.append(("refit", np.array(x), np.array(y),
                               np.array(out)))
            return out

        models.robust_polyfit = rp
        self._interp = rc.interpolate

        class Spl(self._interp.RectBivariateSpline):
            def __call__(s, *a, **k):
                v = super().__call__(*a, **k)
                rec.events.append(("spline", float(np.sum(v))))
                return v

        proxy = types.SimpleNamespace(
            RectBivariateSpline=Spl, interp1d=self._interp.interp1d)
        rc.interpolate = proxy

    def remove(self):
        np.polynomial.polynomial.polyfit = self._polyfit
        np.polynomial.legendre.legfit, np.polynomial.chebyshev.chebfit = self._legfit, self._chebfit
        Calibrator._match_bijective = self._mb
        models.robust_polyfit = self._rp
        rc.interpolate = self._interp


def _trace_arrays(rec, c, s_total):
    """Turn the event stream into per-draw arrays."""
    peaks = rec.ransac_peaks
    cand = [np.asarray(rec.cand_dict[p]) for p in peaks]
    s = c.sample_size
    draws = []
    cur = None
    n_abort = 0
    for ev in rec.events:
        kind = ev[0]
        if kind == "fit":
            cur = dict(x=ev[1][:s], y=ev[2][:s], coeffs=ev[3], status=None,
                       cost=np.nan, n_inl=0, n_match=0, accepted=False,
                       refit=None, err_sum=np.nan, weight=np.nan)
            draws.append(cur)
        elif kind == "log":
            m = ev[1]
            if m.startswith("Intercept exceeds"):
                cur["status"] = ST_INTERCEPT
            elif m.startswith("Solution is not monoton"):
                cur["status"] = ST_MONOTONIC
            elif m.startswith("Not possible to draw"):
                n_abort += 1
            elif (m.startswith("Too few good") or m.startswith("Fit error too")
                  or m.startswith("Not enough matched")):
                cur["accepted"] = False
                cur["rejected_reason"] = m
        elif kind == "match":
            err, mx, my = ev[1], ev[2], ev[3]
            if cur is None:  # the initial cost of a pre-existing fit (calibrator.py:512-515), before any draw
                continue
            cur["status"] = ST_SCORED if len(mx) else ST_EMPTY
            e = np.minimum(err, c.ransac_tolerance)
            cur["err_clipped"] = e
            cur["n_match"] = len(e)
            cur["n_inl"] = int(np.sum(e < c.ransac_tolerance))
            cur["err_sum"] = sum(e)
        elif kind == "spline":
            cur["weight"] = c.hough_weight * ev[1]
            cur["cost"] = (cur["err_sum"] / (cur["n_match"] - len(cur["coeffs"]) + 1)
                           / (cur["weight"] + 1e-9))
        elif kind == "refit":
            cur["refit"] = ev[3]
            cur["refit_x"], cur["refit_y"] = ev[1], ev[2]
            cur["accepted"] = "rejected_reason" not in cur
    # a refit followed by a rejection message: fix flags
    for d in draws:
        if "rejected_reason" in d:
            d["accepted"] = False
    n = len(draws)
    x_idx = np.zeros((n, s), np.int32)
    y_idx = np.zeros((n, s), np.int32)
    for i, d in enumerate(draws):
        for k in range(s):
            pi = int(np.flatnonzero(peaks == d["x"][k])[0])
            x_idx[i, k] = pi
            y_idx[i, k] = int(np.flatnonzero(cand[pi] == d["y"][k])[0])
    out = dict(
        x_idx=x_idx, y_idx=y_idx,
        coeffs=np.array([d["coeffs"] for d in draws]),
        status=np.array([d["status"] for d in draws], np.int32),
        cost=np.array([d["cost"] for d in draws]),
        weight=np.array([d["weight"] for d in draws]),
        n_match=np.array([d["n_match"] for d in draws], np.int32),
        n_inl=np.array([d["n_inl"] for d in draws], np.int32),
        accepted=np.array([d["accepted"] for d in draws]),
        refit_tried=np.array([d["refit"] is not None for d in draws]),
        n_abort=np.int32(n_abort),
        ransac_peaks=peaks,
        cand_len=np.array([len(a) for a in cand], np.int32),
        cand_flat=np.concatenate(cand),
    )
    acc = [d for d in draws if d["accepted"]]
    if acc:
        out["last_refit_x"] = acc[-1]["refit_x"]
        out["last_refit_y"] = acc[-1]["refit_y"]
        out["last_refit_coeffs"] = acc[-1]["refit"]
    return out


def record_case(name, peaks, lines, *, spectrum=None, calib=None, hough=None,
                ransac=None, fit=None, seed=0, atlas_kw=None, known=None):
    calib, hough, ransac, fit = (dict(calib or {}), dict(hough or {}),
                                 dict(ransac or {}), dict(fit or {}))
    cfg = dict(calib=calib, hough=hough, ransac=ransac, fit=fit, seed=seed,
               atlas_kw=atlas_kw or {})
    if known is not None:  # set_known_pairs (calibrator.py:1507-1550): appended to every sample (:568-573)
        cfg["known"] = [list(map(float, known[0])), list(map(float, known[1]))]

    def build():
        c = Calibrator(peaks, spectrum=spectrum)
        c.set_calibrator_properties(seed=seed, **calib)
        c.set_hough_properties(**hough)
        c.set_ransac_properties(**ransac)
        a = Atlas()
        a.add_user_atlas(elements=["X"] * len(lines), wavelengths=lines)
        c.set_atlas(a, **(atlas_kw or {}))
        if known is not None:
            c.set_known_pairs(known[0], known[1])
        return c

    # (1) the reference exactly as shipped: its own (unstable) ranking
    _CANON["on"] = False
    c = build()
    c.do_hough_transform()
    ref_order = c.ht.ref_order.astype(np.int32)
    hp = np.ascontiguousarray(c.hough_points)
    out = dict(
        peaks=np.asarray(peaks, float), lines=np.asarray(lines, float),
        pixel_list=np.asarray(c.pixel_list, float),
        pairs=np.asarray(c.pairs, float),
        limits=np.array([c.min_slope, c.max_slope, c.min_intercept,
                         c.max_intercept]),
        n_points=np.int64(len(hp)),
        points_sha256=np.frombuffer(hashlib.sha256(hp.tobytes()).digest(),
                                    np.uint8),
        points_head=hp[:64].copy(), points_tail=hp[-64:].copy(),
        hist=c.ht.hist.astype(np.int32), xedges=c.ht.xedges,
        yedges=c.ht.yedges, ref_order=ref_order,
    )
    # (2) same reference code, canonical tie order for the ranking
    _CANON["on"] = True
    c = build()
    c.do_hough_transform()
    rec = Recorder()
    rec.install(c)
    try:
        res = c.fit(progress=False, **fit)
    finally:
        rec.remove()
        _CANON["on"] = False
    out["candidate_peak"] = np.array(c.candidate_peak, float)
    out["candidate_arc"] = np.array(c.candidate_arc, float)
    out["n_cand_per_line"] = np.array([len(t[0]) for t in c.candidates],
                                      np.int32)
    out.update({"tr_" + k: v for k, v in
                _trace_arrays(rec, c, c.sample_size).items()})
    out["sample_size"] = np.int32(c.sample_size)
    out["fit_coeff"] = np.array(res[0])
    out["matched_peaks"] = np.array(res[1], float)
    out["matched_atlas"] = np.array(res[2], float)
    out["rms"] = np.float64(res[3])
    out["residual"] = np.array(res[4])
    out["peak_utilisation"] = np.float64(res[5])
    out["atlas_utilisation"] = np.float64(res[6])
    out["config_json"] = np.array(json.dumps(cfg))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    st = out["tr_status"]
    print(f"{name}: P={len(c.pairs)} N={out['n_points']} draws={len(st)} "
          f"icpt={np.sum(st == ST_INTERCEPT)} mono={np.sum(st == ST_MONOTONIC)} "
          f"scored={np.sum(st == ST_SCORED)} accepted={out['tr_accepted'].sum()} "
          f"rms={out['rms']:.4g} -> {os.path.getsize(path) / 1024:.0f} KiB")


def c3_inputs():
    """SURVEY.md §8d 'C3 concrete input'."""
    rng = np.random.default_rng(1234)
    npix = 4096
    co = [4000, 1.2, 2e-5, -3e-9, 1e-13]
    lam = lambda x: np.polynomial.polynomial.polyval(x, co)  # noqa: E731
    px = np.sort(rng.uniform(50, npix - 50, 60))
    peaks = px + rng.normal(0, 0.05, 60)
    atlas = np.sort(np.concatenate(
        [lam(px), rng.uniform(lam(0) - 400, lam(4095) + 400, 240)]))
    return peaks, atlas, npix, float(lam(0)), float(lam(4095))


def polyfit_test_inputs():
    """test/test_polynomial_fit.py:5-15."""
    rs = np.random.RandomState(0)
    peaks = np.sort(rs.random_sample(31) * 1000.0)
    m = np.isclose(peaks[:-1], peaks[1:], atol=5.0)
    m = np.insert(m, 0, False)
    peaks = peaks[~m]
    return peaks, 3000.0 + 5.0 * peaks, 3000.0 + 4 * peaks + 1.0e-3 * peaks ** 2.0


def main():
    peaks, atlas, npix, l0, l1 = c3_inputs()
    record_case(
        "c3_deg4", peaks, atlas, spectrum=np.zeros(npix),
        hough=dict(num_slopes=2000, xbins=100, ybins=100, min_wavelength=l0,
                   max_wavelength=l1, range_tolerance=500,
                   linearity_tolerance=100),
        ransac=dict(sample_size=5, top_n_candidate=5, ransac_tolerance=5),
        fit=dict(max_tries=300, fit_deg=4, candidate_tolerance=10.0),
        seed=3, atlas_kw=dict(candidate_tolerance=10.0))

    pk, wl, wq = polyfit_test_inputs()
    hough = dict(num_slopes=1000, range_tolerance=500.0, xbins=200, ybins=200,
                 min_wavelength=3000.0, max_wavelength=8000.0)
    record_case(
        "poly_linear_deg1", pk, wl, calib=dict(num_pix=1000), hough=hough,
        ransac=dict(minimum_matches=20, minimum_fit_error=1e-12),
        fit=dict(max_tries=200, fit_deg=1), seed=0)
    record_case(
        "poly_quadratic_deg2", pk, wq, calib=dict(num_pix=1000), hough=hough,
        ransac=dict(minimum_matches=20, minimum_fit_error=1e-12),
        fit=dict(max_tries=200, fit_deg=2, candidate_tolerance=5.0), seed=0)

    # fine binning, degree 5, larger sample than deg+1 (over-determined fits),
    # filter_close on: the C2/C5-shaped corner cases at a size
    # that keeps the fixture small
    rng = np.random.default_rng(77)
    npix = 2048
    co = [6400.0, 1.9, -4e-5, 2.5e-8, -4e-12, 3e-16]
    lam = lambda x: np.polynomial.polynomial.polyval(x, co)  # noqa: E731
    px = np.sort(rng.uniform(20, npix - 20, 45))
    peaks = px + rng.normal(0, 0.1, 45)
    lines = np.sort(np.concatenate(
        [lam(px[:38]), rng.uniform(lam(0) - 300, lam(npix - 1) + 300, 70)]))
    record_case(
        "fine_deg5", peaks, lines, spectrum=np.zeros(npix),
        hough=dict(num_slopes=1500, xbins=300, ybins=250,
                   min_wavelength=float(lam(0)) + 60.0,
                   max_wavelength=float(lam(npix - 1)) - 80.0,
                   range_tolerance=400, linearity_tolerance=80),
        ransac=dict(sample_size=7, top_n_candidate=4, ransac_tolerance=4,
                    filter_close=True,
                    minimum_matches=8, minimum_peak_utilisation=0.2),
        fit=dict(max_tries=150, fit_deg=5, candidate_tolerance=6.0), seed=11)


if __name__ == "__main__":
    main()
