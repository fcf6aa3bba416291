"""Host-side logic of the product (no GPU): K2 layout, the RANSAC set-up tables
(a7, calibrator.py:441-523) against the oracle and the golden recordings, the
facade's setters, and the rank-merge rule used by the multi-GPU path."""
import numpy as np
import pytest

from oracle import rascal_oracle as O
from tests.helpers import problem_from_golden


def _setup(g, **over):
    from tests.test_gpu_ransac_refit import setup_from_golden

    return setup_from_golden(g, **over)


def test_ransac_setup_tables(golden):
    g = golden
    s = _setup(g)
    p = problem_from_golden(g)
    assert np.array_equal(s.peaks, p.peaks)
    assert np.array_equal(s.cand_wave, np.concatenate(p.cands))
    assert np.array_equal(np.diff(s.cand_off), [len(c) for c in p.cands])
    # the normalising sum is sequential here (so that rb2_ransac_prepare reproduces the table on the
    # device bit for bit), np.sum's pairwise order in the reference: last-bit differences only
    assert np.allclose(s.prob, p.prob, rtol=1e-14, atol=0)
    assert (s.pix_min, s.pix_max) == (p.pix_min, p.pix_max)
    assert np.array_equal(s.uwaves[s.cand_wid], s.cand_wave)
    # CDF table reproduces searchsorted(cdf, u, 'right') of the sampling law
    cdf = np.cumsum(p.prob)
    u = np.random.default_rng(0).integers(0, 2 ** 32, 20000, dtype=np.uint64)
    idx_dev = np.minimum(np.searchsorted(s.cdf32.astype(np.uint64), u, side="right"), len(cdf) - 1)
    idx_ref = np.minimum(np.searchsorted(cdf / cdf[-1], u / 2.0 ** 32, side="right"), len(cdf) - 1)
    assert np.mean(idx_dev != idx_ref) < 1e-3


def test_weight_tables_reproduce_the_reference_spline(golden):
    """h_lo/h_hi and the pp-form section == RectBivariateSpline(t, ic_min) clamped."""
    g = golden
    s = _setup(g)
    p = problem_from_golden(g)
    t = np.linspace(s.gc_min, s.gc_max, 777)
    ref = p.spline(t, np.full_like(t, 1.0), grid=False)  # gradient ~1 << ic_min: clamps to the first row
    k = np.clip(np.searchsorted(s.pp_breaks, t, side="right") - 1, 0, len(s.pp_breaks) - 2)
    dt = t - s.pp_breaks[k]
    c = s.pp_coef.reshape(-1, 4)[k]
    mine = ((c[:, 3] * dt + c[:, 2]) * dt + c[:, 1]) * dt + c[:, 0]
    assert np.allclose(mine, ref, rtol=1e-10, atol=1e-9)
    if g["hist"].size <= 65536:  # the reference's own spline object evaluated at the clamped corners
        assert s.h_lo == float(p.spline(-1e9, 1.0, grid=False))
        assert s.h_hi == float(p.spline(1e9, 1.0, grid=False))
    else:  # fine binnings: the histogram corners themselves (the interpolating spline reproduces them to rounding)
        assert (s.h_lo, s.h_hi) == (float(g["hist"][0, 0]), float(g["hist"][-1, 0]))
        assert np.isclose(s.h_lo, float(p.spline(-1e9, 1.0, grid=False)), rtol=1e-13, atol=1e-13)
        assert np.isclose(s.h_hi, float(p.spline(1e9, 1.0, grid=False)), rtol=1e-13, atol=1e-13)


def test_filter_close_matches_reference_rule():
    from rascal_b200 import engine

    x = np.array([1, 1, 2, 2, 3, 3, 4.0])
    y = np.array([10, 50, 12, 90, 50, 130, 131.0])
    hist = np.ones((4, 4))
    e = np.linspace(0, 1, 5)
    kw = dict(fit_deg=1, sample_size=2, ransac_tolerance=1.0, min_intercept=0, max_intercept=1,
              pixel_list=np.arange(8), hist=hist, xedges=e, yedges=e + 3000)
    s = engine.RansacSetup(x, y, filter_close=True, **kw)
    p = O.RansacProblem(x, y, filter_close=True, **kw)
    assert np.array_equal(s.x, p.x) and np.array_equal(s.y, p.y)


def test_vote_layout_roundtrip():
    from rascal_b200 import engine

    rng = np.random.default_rng(1)
    peaks = rng.permutation(np.r_[rng.uniform(0, 100, 7), 5.0, 5.0])  # unsorted, duplicated
    lines = rng.uniform(4000, 5000, 11)
    pairs = O.make_pairs(peaks, lines)
    up, off, order, y, wid, uw = engine.vote_layout(pairs)
    assert np.array_equal(up, np.unique(peaks))
    assert off[-1] == len(pairs)
    for u in range(len(up)):
        sel = np.flatnonzero(pairs[:, 0] == up[u])
        assert np.array_equal(order[off[u]:off[u + 1]], sel)  # original order inside a peak
    assert np.array_equal(uw[wid], y) and np.array_equal(y, pairs[order, 1])


def test_facade_setters_follow_reference_defaults():
    from rascal_b200.calibrator import Calibrator, LineList

    c = Calibrator(np.array([10.0, 500.0, 900.0]))
    assert c.num_pix == 1.1 * 900.0 and len(c.pixel_list) == len(np.arange(1.1 * 900.0))
    assert (c.num_slopes, c.xbins, c.ybins) == (2000, 100, 100)
    assert (c.min_wavelength, c.max_wavelength, c.range_tolerance, c.linearity_tolerance) == (3000.0, 9000.0, 500, 100)
    assert (c.sample_size, c.top_n_candidate, c.linear, c.filter_close) == (5, 5, True, False)
    assert (c.ransac_tolerance, c.candidate_weighted, c.hough_weight) == (5, True, 1.0)
    assert (c.minimum_matches, c.minimum_peak_utilisation, c.minimum_fit_error) == (0, 0, 1e-4)
    lim = O.hough_limits(3000.0, 9000.0, 500, 100, c.pixel_list)
    assert (c.min_slope, c.max_slope, c.min_intercept, c.max_intercept) == lim
    # None keeps the previous value; num_pix does not regenerate pixel_list
    c.set_hough_properties(xbins=50)
    assert (c.xbins, c.ybins) == (50, 100)
    c.set_calibrator_properties(num_pix=2000)
    assert c.num_pix == 2000 and len(c.pixel_list) == len(np.arange(1.1 * 900.0))
    c.set_atlas(LineList([4000.0, 5000.0]))
    assert np.array_equal(c.pairs, O.make_pairs(c.peaks, [4000.0, 5000.0]))
    with pytest.raises(AssertionError):
        c.set_ransac_properties(minimum_matches=0)
    with pytest.raises(AssertionError):
        c.set_ransac_properties(minimum_peak_utilisation=1.5)


def test_prior_cost_is_the_reference_initial_cost(golden):
    """calibrator.py:512-515: best_cost = sum(_match_bijective(candidates, peaks, fit_coeff)[0])."""
    g = golden
    s = _setup(g)
    p = problem_from_golden(g)
    for coeffs in (g["fit_coeff"], g["tr_coeffs"][0]):
        err, _, _ = O.match_bijective(p, coeffs)
        assert s.prior_cost(coeffs) == float(sum(err))


def test_degenerate_sampling_law_raises_like_the_reference():
    """A peak exactly on an interior edge of the 10-bin histogram whose left bin is empty: the reference's
    weights hold 1/0, np.random.choice raises ValueError('probabilities contain NaN') (calibrator.py:521-540)."""
    import pytest

    from rascal_b200 import engine

    with pytest.raises(ValueError, match="probabilities contain NaN"):
        engine.sampling_cdf32(np.array([100.0, 150.0, 500.0, 700.0, 1100.0]))
