"""The peak-finding / peak-refinement oracle (oracle/peaks.py) against recordings of the
reference's util.refine_peaks (tests/golden/peaks_*.npz) and against SciPy itself (the
third-party code the reference calls: scipy.signal.find_peaks, MINPACK lmdif through
scipy.optimize.leastsq)."""
import glob
import os

import numpy as np
import pytest

from oracle import peaks as OP

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "peaks_*.npz")))


def _kw(g):
    return {k: (None if np.isnan(g["fp_" + k]) else float(g["fp_" + k])) for k in ("height", "distance", "prominence")}


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[6:-4] for p in GOLD])
def test_find_peaks_matches_recording(path):
    g = np.load(path)
    peaks, props = OP.find_peaks(g["spectrum"], **_kw(g))
    assert np.array_equal(peaks, g["peaks"])
    if "prominences" in g.files:
        assert np.array_equal(props["prominences"], g["prominences"])


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[6:-4] for p in GOLD])
def test_refine_peaks_matches_recording(path):
    g = np.load(path)
    out = OP.refine_peaks(g["spectrum"], g["peaks"], window_width=int(g["window_width"]))
    # the restated lmdif walks the same iterates as MINPACK: bit-identical in practice
    assert out.shape == g["refined"].shape
    assert np.max(np.abs(out - g["refined"])) <= 1e-10


def test_find_peaks_matches_scipy_random():
    from scipy.signal import find_peaks

    rng = np.random.default_rng(5)
    for t in range(60):
        n = int(rng.integers(3, 400))
        x = rng.normal(0, 1, n).cumsum() if t % 2 else rng.normal(0, 1, n)
        if t % 3 == 0:
            x = np.round(x * 2) / 2  # plateaus and ties
        kw = dict(height=[None, 0.0, 0.5][t % 3], distance=[None, 1, 3.5, 10][t % 4], prominence=[None, 0.3, 1.0][t % 3])
        a, pa = find_peaks(x, **kw)
        b, pb = OP.find_peaks(x, **kw)
        if kw["distance"] is not None and len(np.unique(x[OP.local_maxima_1d(x)])) != len(OP.local_maxima_1d(x)):
            continue  # SciPy's priority order among EQUAL heights is an unstable argsort: unspecified
        assert np.array_equal(a, b), (t, kw)
        if kw["prominence"] is not None:
            assert np.array_equal(pa["prominences"], pb["prominences"])


def test_lmdif_matches_minpack():
    from scipy.optimize import leastsq

    rng = np.random.default_rng(1)
    for t in range(120):
        m = int(rng.integers(4, 21))
        x = np.arange(m, dtype=float)
        c, s = rng.uniform(1, m - 1), rng.uniform(0.6, 3)
        y = np.exp(-(x - c) ** 2 / (2 * s * s)) + rng.normal(0, 0.05 if t % 3 else 0.3, m)
        if t % 7 == 0:
            y += 0.6 * np.exp(-(x - c - 3) ** 2 / 2)
        y /= y.max()
        mean = np.sum(x * y) / m
        p0 = [1.0, mean, np.sum(y * (x - mean) ** 2) / m]

        def f(p):
            return OP.gauss(x, *p) - y

        r = leastsq(f, p0, full_output=1)
        xo, ier, nfev = OP.lmdif(f, p0)
        assert ier == r[4] and nfev == r[2]["nfev"]
        assert np.max(np.abs(xo - r[0])) <= 1e-12 * max(1.0, np.max(np.abs(r[0])))


def test_refine_peaks_lmdif_equals_curve_fit():
    g = np.load(GOLD[0])
    a = OP.refine_peaks(g["spectrum"], g["peaks"], window_width=int(g["window_width"]), fitter="scipy")
    b = OP.refine_peaks(g["spectrum"], g["peaks"], window_width=int(g["window_width"]), fitter="lmdif")
    assert np.array_equal(a, b)
