"""The oracle (oracle/rascal_oracle.py) against recordings of the real reference.

These tests pin the oracle: every deterministic intermediate of the reference
(Hough counts, edges, candidates, per-draw coefficients/status/cost, seeded
sample stream, final model) must be reproduced bit for bit.
"""
import hashlib

import numpy as np
import pytest

from oracle import rascal_oracle as O
from tests.helpers import oracle_from_golden, problem_from_golden


def test_limits_and_pairs(golden):
    g = golden
    h = g["config"]["hough"]
    lim = O.hough_limits(h.get("min_wavelength", 3000.0), h.get("max_wavelength", 9000.0),
                         h.get("range_tolerance", 500), h.get("linearity_tolerance", 100),
                         g["pixel_list"])
    assert np.array_equal(np.array(lim), g["limits"])
    assert np.array_equal(O.make_pairs(g["peaks"], g["lines"]), g["pairs"])


def test_hough_points_and_histogram(golden):
    g = golden
    h = g["config"]["hough"]
    lo, hi, min_i, max_i = g["limits"]
    pts = O.hough_points(g["pairs"][:, 0], g["pairs"][:, 1], lo, hi,
                         int(h["num_slopes"]), min_i, max_i)
    assert len(pts) == int(g["n_points"])
    digest = hashlib.sha256(np.ascontiguousarray(pts).tobytes()).digest()
    assert digest == g["points_sha256"].tobytes()
    hist, xe, ye = O.hough_histogram(pts, h["xbins"], h["ybins"])
    assert np.array_equal(hist, g["hist"])
    assert np.array_equal(xe, g["xedges"]) and np.array_equal(ye, g["yedges"])


def test_reference_order_is_a_valid_ranking(golden):
    """The reference's own (unstable) order and the canonical order agree on the
    count sequence; they may differ only inside runs of equal counts."""
    g = golden
    flat = g["hist"].ravel()
    canon = O.rank_bins(g["hist"].astype(float))
    assert np.array_equal(flat[g["ref_order"]], flat[canon])
    assert np.array_equal(np.sort(g["ref_order"]), np.arange(flat.size))


def test_candidates(golden):
    g = golden
    r = oracle_from_golden(g, max_tries=0)
    assert np.array_equal(np.array(r["candidate_peak"]), g["candidate_peak"])
    assert np.array_equal(np.array(r["candidate_arc"]), g["candidate_arc"])


def test_replayed_hypotheses(golden):
    """Every drawn sample of the recorded run, re-evaluated by the oracle."""
    g = golden
    prob = problem_from_golden(g)
    assert np.array_equal(prob.peaks, g["tr_ransac_peaks"])
    assert np.array_equal(np.concatenate(prob.cands), g["tr_cand_flat"])
    n = len(g["tr_status"])
    step = max(1, n // 1500)  # the monotonic grid makes each call ~0.3 ms
    for i in list(range(0, n, step)) + list(np.flatnonzero(g["tr_status"] == O.ST_SCORED)):
        xi, yi = g["tr_x_idx"][i], g["tr_y_idx"][i]
        r = O.evaluate_hypothesis(prob, prob.peaks[xi],
                                  [prob.cands[k][j] for k, j in zip(xi, yi)])
        assert np.array_equal(r["coeffs"], g["tr_coeffs"][i])
        assert r["status"] == g["tr_status"][i]
        if r["status"] == O.ST_SCORED:
            assert r["cost"] == g["tr_cost"][i]
            assert r["n_inliers"] == g["tr_n_inl"][i]
            assert r["n_match"] == g["tr_n_match"][i]


def test_seeded_loop_reproduces_reference(golden, request):
    g = golden
    if request.node.callspec.id == "c5_deimos_full":
        # the whole path again at 10^6 Hough lines costs the oracle 45 s for a 40-try loop; its candidate
        # generation at that size is pinned by test_candidates, its loop by the same arc at 400 x 400 (c5_deimos)
        pytest.skip("covered by test_candidates[c5_deimos_full] + test_seeded_loop_reproduces_reference[c5_deimos]")
    trace = []
    r = oracle_from_golden(g, trace=trace)
    draws = [t for t in trace if t["status"] != O.ST_ABORT]
    assert len(draws) == len(g["tr_status"])
    assert np.array_equal(np.array([t["x_idx"] for t in draws]), g["tr_x_idx"])
    assert np.array_equal(np.array([t["y_idx"] for t in draws]), g["tr_y_idx"])
    assert np.array_equal(np.array([t["accepted"] for t in draws]), g["tr_accepted"])
    b = r["best"]
    assert np.array_equal(b["best_p"], g["fit_coeff"])
    assert np.array_equal(np.array(b["matched_peaks"]), g["matched_peaks"])
    assert np.array_equal(np.array(b["matched_atlas"]), g["matched_atlas"])
    assert b["rms"] == g["rms"]
    assert np.array_equal(b["residual"], g["residual"])


def test_order_independent_rule(golden):
    """result == min-cost acceptable hypothesis, latest among ties."""
    g = golden
    prob = problem_from_golden(g)
    scored = np.flatnonzero(g["tr_status"] == O.ST_SCORED)
    res = []
    for i in scored:
        xi, yi = g["tr_x_idx"][i], g["tr_y_idx"][i]
        res.append(O.evaluate_hypothesis(prob, prob.peaks[xi],
                                         [prob.cands[k][j] for k, j in zip(xi, yi)]))
    k, acc = O.best_of(prob, res)
    assert scored[k] == np.flatnonzero(g["tr_accepted"])[-1]
    assert np.array_equal(acc[1], g["fit_coeff"])


def test_exact_refit_close_to_trf(golden):
    """Evidence for DESIGN.md: SciPy's TRF result vs the exact Huber/LSQ optimum
    on the same inlier set - predictions agree far below 1e-6 A even where the
    coefficients only agree to ~1e-5 relative."""
    g = golden
    x, y = g["tr_last_refit_x"], g["tr_last_refit_y"]
    deg = len(g["fit_coeff"]) - 1
    ce = O.exact_refit(x, y, deg)
    P = np.polynomial.polynomial
    assert np.max(np.abs(P.polyval(x, ce) - P.polyval(x, g["tr_last_refit_coeffs"]))) < 1e-6
