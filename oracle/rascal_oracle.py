"""CPU oracle for the rascal calibration hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain NumPy/SciPy, the arithmetic of the reference
(jveitchmichaelis/rascal v0.3.9) for the one path this repo accelerates:

    Hough accumulation -> candidate vote -> RANSAC fit/score -> winner refit.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; the product package ``rascal_b200``
never does (it fails loudly when the CUDA library is missing).

Pinning: ``tests/golden/make_golden.py`` runs the *imported* reference in the
build container with recording hooks and commits the inputs/outputs under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
here against those recordings (bit-exact where the reference is
deterministic).  Third-party arithmetic (``np.histogram2d``, ``polyfit``,
``scipy.optimize.least_squares``, ``RectBivariateSpline``, legacy
``RandomState``) is *called*, not restated, exactly as the reference calls it.

Every function cites the reference lines (``src/rascal/...``) it follows.
Deliberate canonicalisation (two places): ties of the reference's two default
(unstable, build- and CPU-dependent) argsorts.  Bin ranking: we use
``argsort(kind='stable')[::-1]`` = descending count, ties by descending flat
index.  Per-peak aggregate ranking in
``most_common_candidates`` (calibrator.py:212): ``argsort(-agg, kind='stable')``;
ties only occur in practice with ``candidate_weighted=False`` (integer counts).
``tests/test_oracle_live_reference.py`` holds this module to the live reference
on new random problems.
"""

from __future__ import annotations

import numpy as np
import scipy.optimize
from scipy import interpolate

P = np.polynomial.polynomial

# ---------------------------------------------------------------------------
# a2: derived Hough limits      calibrator.py:1088-1116
# ---------------------------------------------------------------------------


def hough_limits(min_wavelength, max_wavelength, range_tolerance,
                 linearity_tolerance, pixel_list):
    """(min_slope, max_slope, min_intercept, max_intercept); calibrator.py:1088-1116."""
    min_i = min_wavelength - range_tolerance
    max_i = min_wavelength + range_tolerance
    span = np.ptp(pixel_list)
    lo = ((max_wavelength - range_tolerance - linearity_tolerance)
          - (min_i + range_tolerance + linearity_tolerance)) / span
    hi = ((max_wavelength + range_tolerance + linearity_tolerance)
          - (min_i - range_tolerance - linearity_tolerance)) / span
    return lo, hi, min_i, max_i


def make_pairs(peaks, lines):
    """Cartesian product, peak-major; calibrator.py:103-133 (no polygon mask)."""
    peaks = np.asarray(peaks, dtype=float)
    lines = np.asarray(lines, dtype=float)
    return np.column_stack((np.repeat(peaks, len(lines)),
                            np.tile(lines, len(peaks))))


# ---------------------------------------------------------------------------
# a3: Hough point generation    houghtransform.py:60-92
# ---------------------------------------------------------------------------


def hough_points(x, y, min_slope, max_slope, num_slopes, min_intercept,
                 max_intercept):
    """[N,2] (gradient, intercept), slope-major; houghtransform.py:78-92."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    slopes = np.linspace(min_slope, max_slope, num_slopes)
    icpt = (y[None, :] - slopes[:, None] * x[None, :])  # == y - np.outer(slopes, x)
    keep = (min_intercept <= icpt) & (icpt <= max_intercept)
    grad = np.broadcast_to(slopes[:, None], icpt.shape)
    return np.column_stack((grad[keep], icpt[keep]))


def hough_points_brute_force(x, y, min_slope, max_slope, min_intercept,
                             max_intercept):
    """Pair-of-pairs lines; houghtransform.py:94-140 (strict inequalities)."""
    order = np.argsort(x)
    x = np.asarray(x, dtype=float)[order]
    y = np.asarray(y, dtype=float)[order]
    gs, cs = [], []
    for i in range(len(x) - 1):
        g = (y[i + 1:] - y[i]) / (x[i + 1:] - x[i])
        gs.append(g)
        cs.append(y[i + 1:] - g * x[i + 1:])
    g = np.concatenate(gs)
    c = np.concatenate(cs)
    keep = ((g > min_slope) & (g < max_slope)
            & (c > min_intercept) & (c < max_intercept))
    return np.column_stack((g[keep], c[keep]))


# ---------------------------------------------------------------------------
# a4: histogram + ranking       houghtransform.py:167-210
# ---------------------------------------------------------------------------


def hough_histogram(points, xbins, ybins):
    """np.histogram2d exactly as houghtransform.py:183-187."""
    return np.histogram2d(points[:, 0], points[:, 1], bins=(xbins, ybins))


def rank_bins(hist):
    """Canonical ranking: count descending, ties by descending flat index.

    houghtransform.py:191-195 uses the default (unstable) argsort; this is the
    one documented canonicalisation.
    """
    return np.argsort(hist.ravel(), kind="stable")[::-1]


def hough_lines(hist, xedges, yedges, order):
    """[B,2] bin-centre lines in rank order; houghtransform.py:197-210."""
    i, j = np.unravel_index(order, hist.shape)
    hx = (xedges[1] - xedges[0]) / 2
    hy = (yedges[1] - yedges[0]) / 2
    return np.column_stack((xedges[i] + hx, yedges[j] + hy))


# ---------------------------------------------------------------------------
# a5: candidate points          calibrator.py:219-261, util.py:426-447
# ---------------------------------------------------------------------------


def candidate_sigma(range_tolerance, linearity_tolerance):
    """calibrator.py:256."""
    return (range_tolerance + linearity_tolerance) * 1.1775


def candidate_points_linear(lines, pairs, candidate_tolerance, sigma):
    """List of (x_sel, lambda_sel, w_sel) per ranked line; calibrator.py:241-261."""
    px = pairs[:, 0]
    actual = pairs[:, 1]
    out = []
    denom = 2 * sigma ** 2 + 1e-9
    for g, c in lines:
        predicted = g * px + c
        diff = np.abs(predicted - actual)
        m = diff <= candidate_tolerance
        w = 1.0 * np.exp(-((actual[m] - predicted[m]) ** 2) / denom)
        out.append((px[m], actual[m], w))
    return out


# ---------------------------------------------------------------------------
# a6: per-peak vote             calibrator.py:154-217   (quirk: SURVEY A.3)
# ---------------------------------------------------------------------------


def candidate_points_poly(peaks, lines, fit_coeff, n_lines, range_tolerance, candidate_tolerance, rs,
                          polyval=P.polyval):
    """linear=False: Calibrator._get_candidate_points_poly, calibrator.py:263-315 (EXPERIMENTAL there).
    `rs`: the legacy RandomState the reference's np.random.random call draws from."""
    peaks = np.asarray(peaks, dtype=float)
    actual = np.array(lines, dtype=float)
    out = []
    delta = rs.random_sample(n_lines) * range_tolerance * 2.0 - range_tolerance
    for d in delta:
        predicted = polyval(peaks, fit_coeff) + d
        diff = np.abs(actual - predicted)
        mask = diff < candidate_tolerance
        if np.sum(mask) > 0:
            weight = np.exp(-((actual[mask] - predicted[mask]) ** 2) / (2 * range_tolerance ** 2 + 1e-9))
            out.append([peaks[mask], actual[mask], weight])
    return out


def adjust_polyfit(delta, fit, peaks, lines, pixel_list, tolerance, min_frac, polyval=P.polyval):
    """Calibrator._adjust_polyfit, calibrator.py:754-820: the objective of match_peaks(refine=True)."""
    fit_new = np.array(fit, dtype=float)
    atlas_lines = np.asarray(lines, dtype=float)
    for i, d in enumerate(delta):
        fit_new[i] += d
    xm, ym = [], []
    for p in peaks:
        x = polyval(p, fit_new)
        diff_abs = np.abs(atlas_lines - x)
        idx = np.argmin(diff_abs)
        if diff_abs[idx] < tolerance:
            xm.append(p)
            ym.append(atlas_lines[idx])
    xm, ym = np.array(xm), np.array(ym)
    dof = len(xm) - len(fit_new) - 1
    if dof < 1:
        return np.inf
    if len(xm) < len(peaks) * min_frac:
        return np.inf
    if not np.all(np.diff(polyval(np.sort(pixel_list), fit_new)) > 0):
        return np.inf
    return np.sum((ym - polyval(xm, fit_new)) ** 2.0) / dof


def most_common_candidates(candidates, top_n, weighted):
    """(candidate_peak, candidate_arc) lists; calibrator.py:174-217.

    Reproduces the indexing quirk: positions from argsort over the *unique*
    wavelengths index the *ordered match list*.  Tie order of the argsort is
    canonicalised to stable.
    """
    pk = np.concatenate([c[0] for c in candidates]) if candidates else np.array([])
    wv = np.concatenate([c[1] for c in candidates]) if candidates else np.array([])
    pr = np.concatenate([c[2] for c in candidates]) if candidates else np.array([])
    out_p, out_w = [], []
    for peak in np.unique(pk):
        sel = pk == peak
        W = wv[sel]
        cnt = pr[sel] if weighted else np.ones_like(pr[sel])
        U = np.unique(W)
        n = int(min(top_n, len(U)))
        agg = np.zeros_like(U)
        for j, w in enumerate(U):
            agg[j] = np.sum(cnt[W == w])
        out_p.extend([peak] * n)
        out_w.extend(W[np.argsort(-agg, kind="stable")[:n]])
    return out_p, out_w


# ---------------------------------------------------------------------------
# a7: RANSAC setup              calibrator.py:441-523
# ---------------------------------------------------------------------------


class RansacProblem:
    """Read-only inputs of the hypothesis loop (what K3 consumes)."""

    def __init__(self, cand_peak, cand_arc, *, fit_deg, sample_size,
                 ransac_tolerance, min_intercept, max_intercept, pixel_list,
                 hist, xedges, yedges, hough_weight=1.0, filter_close=False,
                 minimum_matches=0, minimum_peak_utilisation=0.0,
                 minimum_fit_error=1e-4, n_peaks_total=None,
                 pix_known=None, wave_known=None, fit_type="poly"):
        # calibrator.py:1621-1636: the basis only swaps the two callables; everything else (the
        # monomial util._derivative, the monomial models.robust_polyfit) stays as it is
        self.fit_type = fit_type
        self.polyfit, self.polyval = BASES[fit_type]
        x = np.array(cand_peak, dtype=float)
        y = np.array(cand_arc, dtype=float)
        if filter_close:  # calibrator.py:457-464
            uy = np.unique(y)
            idx = np.argwhere(uy[1:] - uy[0:-1] < 3 * ransac_tolerance)
            keep = np.argwhere((y == uy[idx]).sum(0) == 0)
            y = y[keep].flatten()
            x = x[keep].flatten()
        self.x, self.y = x, y
        self.fit_deg = int(fit_deg)
        self.sample_size = int(sample_size)
        self.tol = ransac_tolerance
        self.min_intercept = min_intercept
        self.max_intercept = max_intercept
        self.pixel_list = np.asarray(pixel_list)
        self.hough_weight = hough_weight
        self.minimum_matches = minimum_matches
        self.minimum_peak_utilisation = minimum_peak_utilisation
        self.minimum_fit_error = minimum_fit_error
        self.pix_known = pix_known
        self.wave_known = wave_known
        self.peaks = np.sort(np.unique(x))  # :488
        self.n_peaks_total = (len(self.peaks) if n_peaks_total is None
                              else n_peaks_total)
        self.cands = [y[x == p] for p in self.peaks]  # :494-495
        self.enough = len(self.peaks) > self.fit_deg  # :468
        # Hough-density spline, :497-506 (x = gradient centres, y = intercept centres)
        hx = (xedges[1] - xedges[0]) / 2.0
        hy = (yedges[1] - yedges[0]) / 2.0
        self.spline = interpolate.RectBivariateSpline(
            xedges[1:] - hx, yedges[1:] - hy, hist)
        # sampling probabilities, :521-523 (index -1 wrap-around quirk kept)
        if self.enough:
            h = np.histogram(self.peaks, bins=10)
            prob = 1.0 / h[0][np.digitize(self.peaks, h[1], right=True) - 1]
            self.prob = prob / np.sum(prob)
        else:
            self.prob = None
        ptp = np.ptp(self.peaks) if len(self.peaks) else 0.0
        self.pix_min = self.peaks[0] - ptp * 0.2 if len(self.peaks) else 0.0  # :585
        self.pix_max = self.peaks[-1] + ptp * 0.2 if len(self.peaks) else 0.0  # :586


# ---------------------------------------------------------------------------
# a9-a13: one hypothesis
# ---------------------------------------------------------------------------

BASES = {"poly": (P.polyfit, P.polyval),
         "legendre": (np.polynomial.legendre.legfit, np.polynomial.legendre.legval),
         "chebyshev": (np.polynomial.chebyshev.chebfit, np.polynomial.chebyshev.chebval)}

ST_ABORT = 0        # impossible draw (try consumed)          :557-566
ST_INTERCEPT = 1    # intercept test failed                   :578-582
ST_MONOTONIC = 2    # not strictly increasing                 :585-600
ST_EMPTY = 3        # no matches                              :607-608
ST_SCORED = 4       # scored (cost valid)


def deriv_coeffs(p):
    """util.py:526-545."""
    return [i * p[i] for i in range(1, len(p))]


def match_bijective(prob: RansacProblem, coeffs):
    """(err, matched_x, matched_y) sorted by wavelength; calibrator.py:341-380."""
    n = len(prob.peaks)
    err = np.empty(n)
    my = np.empty(n)
    for k, peak in enumerate(prob.peaks):
        fit = prob.polyval(peak, coeffs)
        e = np.abs(fit - prob.cands[k])
        i = np.argmin(e)
        err[k] = e[i]
        my[k] = prob.cands[k][i]
    fx, fy, fe = [], [], []
    for w in np.unique(my):
        m = my == w
        i = np.argmin(err[m])
        fy.append(w)
        fx.append(prob.peaks[m][i])
        fe.append(err[m][i])
    return np.array(fe), np.array(fx), np.array(fy)


def hough_weight_sum(prob: RansacProblem, coeffs):
    """hough_weight * sum(spline(intercept, gradient)); calibrator.py:614-627."""
    px = prob.pixel_list
    wave = prob.polyval(px, coeffs)
    grad = prob.polyval(px, deriv_coeffs(coeffs))
    icpt = wave - grad * px
    return prob.hough_weight * np.sum(prob.spline(icpt, grad, grid=False))


def is_monotonic(prob: RansacProblem, coeffs):
    """calibrator.py:585-596."""
    grid = np.arange(prob.pix_min, prob.pix_max, 1)
    return bool(np.all(np.diff(prob.polyval(grid, coeffs)) > 0))


def evaluate_hypothesis(prob: RansacProblem, x_hat, y_hat):
    """Fit + tests + score for one drawn sample; calibrator.py:569-633.

    Returns dict(status, coeffs, cost, n_match, n_inliers, err, mx, my).
    """
    x_hat = np.asarray(x_hat, dtype=float)
    y_hat = np.asarray(y_hat, dtype=float)
    if prob.pix_known is not None:
        x_hat = np.concatenate((x_hat, prob.pix_known))
        y_hat = np.concatenate((y_hat, prob.wave_known))
    c = prob.polyfit(x_hat, y_hat, prob.fit_deg)
    r = dict(status=ST_SCORED, coeffs=c, cost=np.nan, n_match=0, n_inliers=0)
    if (c[0] < prob.min_intercept) | (c[0] > prob.max_intercept):
        r["status"] = ST_INTERCEPT
        return r
    if not is_monotonic(prob, c):
        r["status"] = ST_MONOTONIC
        return r
    err, mx, my = match_bijective(prob, c)
    if len(mx) == 0:
        r["status"] = ST_EMPTY
        return r
    err[err > prob.tol] = prob.tol
    weight = hough_weight_sum(prob, c)
    cost = sum(err) / (len(err) - len(c) + 1) / (weight + 1e-9)
    inl = err < prob.tol
    r.update(cost=cost, n_match=len(err), n_inliers=int(sum(inl)), err=err,
             mx=mx, my=my, weight=weight)
    return r


# ---------------------------------------------------------------------------
# a15: robust refit             models.py:10-123
# ---------------------------------------------------------------------------


def _residual_fn(a, x, y, degree):
    """models.py:10-54: y - (a[-1] + sum_i a[degree-i] * x**i), same op order."""
    t = a[-1]
    for i in range(1, int(degree + 1)):
        t += a[int(degree - i)] * x ** i
    return y - t


def robust_polyfit(x, y, degree, info=None):
    """SciPy TRF/Huber on std-normalised data; models.py:70-123 (x0=None).
    `info` (a dict) receives SciPy's nfev and termination status."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    xs, ys = x.std(), y.std()
    res = scipy.optimize.least_squares(
        _residual_fn, np.ones(degree + 1), args=(x / xs, y / ys, degree),
        loss="huber", diff_step=1e-5)
    if info is not None:
        info.update(nfev=res.nfev, status=res.status)
    p = res.x
    p *= y.std()
    for i in range(0, degree):
        p[i] /= x.std() ** (degree - i)
    return p[::-1]


def exact_refit(x, y, degree):
    """The fixed point robust_polyfit iterates towards when every normalised
    residual is inside Huber's quadratic zone (|r| <= 1): ordinary least squares
    on (x/std, y/std).  Low-order-first coefficients in raw units."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    xs, ys = x.std(), y.std()
    V = np.vander(x / xs, degree + 1, increasing=True)
    c, *_ = np.linalg.lstsq(V, y / ys, rcond=None)
    return c * ys / xs ** np.arange(degree + 1)



def match_peaks(peaks, lines, fit_coeff, tolerance=10.0, fit_deg=None,
                robust_refit=True, refit=None):
    """Calibrator.match_peaks with refine=False (calibrator.py:1824-1956).

    The reference's permutation search (:1841-1876) only completes when every
    peak has at most one atlas line inside ``tolerance``: with an ambiguous
    peak it raises (TypeError at :1867 under numpy >= 1.24; with that index
    fixed, np.array() of a ragged list / a length mismatch in polyfit).  For
    such a peak this restatement takes the NEAREST line (first on ties) and
    counts it in ``n_ambiguous``; with n_ambiguous == 0 the result is the
    reference's.  Returns a dict (fit_coeff, matched_peaks, matched_atlas,
    rms, residuals, polyfit_coeff, polyfit_err, polyfit_rms, n_ambiguous)."""
    polyfit = np.polynomial.polynomial.polyfit
    polyval = np.polynomial.polynomial.polyval
    fit_coeff = np.asarray(fit_coeff, dtype=np.float64)
    if fit_deg is None:
        fit_deg = len(fit_coeff) - 1  # :1795-1796
    atlas_lines = np.asarray(lines, dtype=np.float64)
    mp, ma, n_amb = [], [], 0
    for p in peaks:  # :1830-1838
        x = polyval(p, fit_coeff)
        diff = atlas_lines - x
        hit = np.abs(diff) < tolerance
        if hit.any():
            idx = np.flatnonzero(hit)
            n_amb += len(idx) > 1
            mp.append(p)
            ma.append(atlas_lines[idx[np.argmin(np.abs(diff[idx]))]])
    mp, ma = np.array(mp, dtype=np.float64), np.array(ma, dtype=np.float64)
    out = dict(matched_peaks=mp, matched_atlas=ma, n_ambiguous=int(n_amb),
               fit_coeff=None, rms=None, residuals=None,
               polyfit_coeff=None, polyfit_err=None, polyfit_rms=None)
    if len(mp) <= fit_deg:
        return out
    pc = polyfit(mp, ma, fit_deg)  # :1893
    res = np.abs(ma - polyval(mp, pc))  # :1895-1897
    out.update(polyfit_coeff=pc, polyfit_err=float(np.sum(res)),
               polyfit_rms=float(np.sqrt(np.nansum(res ** 2.0) / len(res))))
    if robust_refit:  # :1915-1941
        c = (refit or robust_polyfit)(mp, ma, fit_deg)
        r = ma - polyval(mp, c)
        out.update(fit_coeff=np.asarray(c), residuals=r,
                   rms=float(np.sqrt(np.nansum(r ** 2.0) / len(r))))
    else:
        out.update(residuals=res, rms=out["polyfit_rms"])
    return out


def acceptance(prob: RansacProblem, r, refit=robust_polyfit):
    """a14 checks after `cost <= best_cost`; calibrator.py:640-700.

    Returns (accepted, coeffs, rms, residual, matched_peaks, matched_atlas).
    """
    mask = r["err"] < prob.tol
    n_inl = int(sum(mask))
    mp, ma = r["mx"][mask], r["my"][mask]
    if len(mp) <= prob.fit_deg:
        return False, None, None, None, mp, ma
    coeffs = refit(mp, ma, prob.fit_deg)
    res = prob.polyval(mp, coeffs) - ma  # the basis' polyval on the MONOMIAL refit coefficients (:664)
    res[np.abs(res) > prob.tol] = prob.tol
    rms = np.sqrt(np.mean(res ** 2))
    if rms < prob.minimum_fit_error:
        return False, coeffs, rms, res, mp, ma
    if n_inl < prob.minimum_matches:
        return False, coeffs, rms, res, mp, ma
    if n_inl < prob.minimum_peak_utilisation * prob.n_peaks_total:
        return False, coeffs, rms, res, mp, ma
    return True, coeffs, rms, res, mp, ma


# ---------------------------------------------------------------------------
# a8 + loop: the sequential reference loop (legacy RNG)   calibrator.py:525-752
# ---------------------------------------------------------------------------


def draw_sample(prob: RansacProblem, rs):
    """One draw with the legacy RandomState call sequence; calibrator.py:538-566.

    Returns (peak_idx[s], cand_idx[s]) or None when the draw is impossible.
    """
    x_hat = rs.choice(prob.peaks, prob.sample_size, replace=False, p=prob.prob)
    y_hat, xi, yi = [], [], []
    for _x in x_hat:
        k = int(np.searchsorted(prob.peaks, _x))
        choice = prob.cands[k]
        if set(choice).issubset(set(y_hat)):
            return None
        y = rs.choice(choice)
        while y in y_hat:
            y = rs.choice(choice)
        y_hat.append(y)
        xi.append(k)
        yi.append(int(np.flatnonzero(choice == y)[0]))
    return np.array(xi), np.array(yi)


def ransac_sequential(prob: RansacProblem, max_tries, rs, fit_coeff=None,
                      refit=robust_polyfit, trace=None):
    """The reference loop, draw for draw; calibrator.py:441-752.

    `trace` (a list) receives one dict per draw: x_idx, y_idx, status, cost,
    n_inliers, accepted.
    Returns dict(best_p, rms, residual, n_inliers, valid, matched_peaks,
    matched_atlas, best_cost).
    """
    best = dict(best_p=None, rms=1e50, residual=None, n_inliers=0, valid=False,
                matched_peaks=[], matched_atlas=[], best_cost=1e50)
    if not prob.enough:
        return best
    if fit_coeff is not None:  # :512-515
        err, _, _ = match_bijective(prob, fit_coeff)
        best["best_cost"] = sum(err)
        best["rms"] = np.sqrt(np.mean(err ** 2.0))
    for _ in range(int(max_tries)):
        while True:
            d = draw_sample(prob, rs)
            if d is None:
                if trace is not None:
                    trace.append(dict(status=ST_ABORT))
                break
            xi, yi = d
            r = evaluate_hypothesis(
                prob, prob.peaks[xi], [prob.cands[k][j] for k, j in zip(xi, yi)])
            r["x_idx"], r["y_idx"], r["accepted"] = xi, yi, False
            if trace is not None:
                trace.append(r)
            if r["status"] != ST_SCORED:
                continue
            if r["cost"] <= best["best_cost"]:
                ok, coeffs, rms, res, mp, ma = acceptance(prob, r, refit)
                if not ok:
                    continue
                r["accepted"] = True
                best.update(best_cost=r["cost"], n_inliers=r["n_inliers"],
                            best_p=coeffs, rms=rms, residual=res,
                            matched_peaks=list(mp), matched_atlas=list(ma))
            break
    best["valid"] = best["n_inliers"] > prob.fit_deg + 1
    return best


def best_of(prob: RansacProblem, results, cost_ceiling=1e50,
            refit=robust_polyfit):
    """Order-independent statement of the loop's outcome: among the
    scored hypotheses, in draw order, the acceptable one of minimal cost, the
    latest among equal costs.  `results` = list of evaluate_hypothesis dicts.
    Returns (index, acceptance tuple) or (None, None)."""
    order = sorted((i for i, r in enumerate(results)
                    if r["status"] == ST_SCORED and r["cost"] <= cost_ceiling),
                   key=lambda i: (results[i]["cost"], -i))
    for i in order:
        acc = acceptance(prob, results[i], refit)
        if acc[0]:
            return i, acc
    return None, None


# ---------------------------------------------------------------------------
# whole path, as Calibrator.do_hough_transform() + fit() run it
# ---------------------------------------------------------------------------


def run_path(peaks, lines, *, num_pix=None, pixel_list=None, num_slopes=2000,
             xbins=100, ybins=100, min_wavelength=3000.0, max_wavelength=9000.0,
             range_tolerance=500, linearity_tolerance=100, sample_size=5,
             top_n_candidate=5, filter_close=False, ransac_tolerance=5,
             candidate_weighted=True, hough_weight=1.0, minimum_matches=0,
             minimum_peak_utilisation=0.0, minimum_fit_error=1e-4,
             max_tries=500, fit_deg=4, candidate_tolerance=2.0, seed=None,
             refit=robust_polyfit, trace=None, timings=None, fit_type="poly",
             pix_known=None, wave_known=None):
    """do_hough_transform() + fit() of the reference (calibrator.py:1425-1452,
    1552-1715) on explicit inputs.  Returns a dict with every intermediate."""
    import time
    peaks = np.asarray(peaks, dtype=float)
    lines = np.asarray(lines, dtype=float)
    if pixel_list is None:
        pixel_list = np.arange(num_pix if num_pix is not None
                               else 1.1 * max(peaks))
    t0 = time.perf_counter()
    lo, hi, min_i, max_i = hough_limits(min_wavelength, max_wavelength,
                                        range_tolerance, linearity_tolerance,
                                        pixel_list)
    pairs = make_pairs(peaks, lines)
    pts = hough_points(pairs[:, 0], pairs[:, 1], lo, hi, int(num_slopes),
                       min_i, max_i)
    hist, xe, ye = hough_histogram(pts, xbins, ybins)
    order = rank_bins(hist)
    hl = hough_lines(hist, xe, ye, order)
    t1 = time.perf_counter()
    sigma = candidate_sigma(range_tolerance, linearity_tolerance)
    cands = candidate_points_linear(hl, pairs, candidate_tolerance, sigma)
    cp, ca = most_common_candidates(cands, top_n_candidate, candidate_weighted)
    t2 = time.perf_counter()
    s = min(sample_size, len(lines))
    if s <= fit_deg:
        s = fit_deg + 1
    prob = RansacProblem(
        cp, ca, fit_deg=fit_deg, sample_size=s,
        ransac_tolerance=ransac_tolerance, min_intercept=min_i,
        max_intercept=max_i, pixel_list=pixel_list, hist=hist, xedges=xe,
        yedges=ye, hough_weight=hough_weight, filter_close=filter_close,
        minimum_matches=min(minimum_matches, len(lines), len(peaks)),
        minimum_peak_utilisation=minimum_peak_utilisation,
        minimum_fit_error=minimum_fit_error, n_peaks_total=len(peaks), fit_type=fit_type,
        pix_known=pix_known, wave_known=wave_known)
    rs = np.random.RandomState(seed)
    best = ransac_sequential(prob, max_tries, rs, refit=refit, trace=trace)
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(hough=t1 - t0, candidates=t2 - t1, ransac=t3 - t2)
    return dict(pairs=pairs, n_points=len(pts), hist=hist, xedges=xe, yedges=ye,
                order=order, lines=hl, candidate_peak=cp, candidate_arc=ca,
                problem=prob, best=best, limits=(lo, hi, min_i, max_i))
