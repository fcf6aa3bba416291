"""oracle/trf.py (the scalar restatement of SciPy's TRF path that K4 implements) against the
real scipy.optimize.least_squares call models.robust_polyfit makes (models.py:108-114)."""
import numpy as np

from oracle import rascal_oracle as O
from oracle import trf

P = np.polynomial.polynomial


def test_restated_trf_follows_scipy_on_golden_inlier_sets(golden):
    g = golden
    x, y = g["tr_last_refit_x"], g["tr_last_refit_y"]
    deg = len(g["fit_coeff"]) - 1
    p, status, nfev = trf.robust_polyfit_restated(x, y, deg)
    info = {}
    ref = O.robust_polyfit(x, y, deg, info=info)
    assert np.array_equal(ref, g["tr_last_refit_coeffs"])  # this SciPy build == the recording
    # same iteration path: equal evaluation counts on the seven first recordings, one more step on one of the
    # C4 inlier sets (the last ftol test falls on the other side of its threshold: FD-Jacobian noise)
    assert abs(nfev - info["nfev"]) <= 1
    # the end point is reproducible only to TRF's own inexactness
    # (a coefficient whose true value is zero - linear data fitted at degree 4 - has no relative scale: it is
    #  held to what it contributes over the detector, 1e-5 A; the curves themselves are compared below)
    atol = 1e-5 / np.max(np.abs(x)) ** np.arange(deg + 1)
    assert np.allclose(p, ref, rtol=2e-5, atol=atol)
    assert np.max(np.abs(P.polyval(x, p) - P.polyval(x, ref))) < 1e-6


def test_restated_trf_random_problems():
    rng = np.random.default_rng(0)
    worst = 0.0
    for k in range(40):
        deg = int(rng.integers(1, 6))
        m = int(rng.integers(deg + 2, 70))
        x = np.sort(rng.uniform(0, 4000, m))
        c = np.array([rng.uniform(3000, 7000), rng.uniform(0.5, 3), rng.uniform(-1e-4, 1e-4),
                      rng.uniform(-1e-8, 1e-8), rng.uniform(-1e-12, 1e-12), rng.uniform(-1e-16, 1e-16)])[: deg + 1]
        y = P.polyval(x, c) + rng.normal(0, rng.choice([0, 0.1, 1.0, 3.0]), m)
        if k % 7 == 0:
            y[rng.integers(0, m)] += rng.uniform(50, 3000)  # gross outlier: Huber's linear zone
        p, status, nfev = trf.robust_polyfit_restated(x, y, deg)
        info = {}
        ref = O.robust_polyfit(x, y, deg, info=info)
        assert abs(nfev - info["nfev"]) <= 3
        worst = max(worst, np.max(np.abs(P.polyval(x, p) - P.polyval(x, ref))))
    assert worst < 1e-5  # incl. outlier-laden and degree-5 problems; TRF stopping noise


def test_jacobi_svd():
    rng = np.random.default_rng(3)
    for n in (2, 5, 8):
        A = rng.normal(size=(40, n)) * 10.0 ** rng.uniform(-3, 3, n)
        U, s, V = trf.jacobi_svd(A)
        assert np.allclose(U * s @ V.T, A, rtol=1e-12, atol=1e-12 * np.abs(A).max())
        assert np.allclose(s, np.linalg.svd(A, compute_uv=False), rtol=1e-12)
        assert np.all(np.diff(s) <= 0)
