"""K5 (rb2_match_peaks) through the C ABI against the oracle and the reference recordings:
identical association, polyfit / refit within the documented tolerances,
ambiguity handling, ragged batches, device-chained coefficients, facade."""
import numpy as np
import pytest

from oracle import rascal_oracle as O
from tests.conftest import MATCH_CASES, load_golden
from tests.test_oracle_match import lines_of, load_match, tol_of

pytestmark = pytest.mark.gpu


def run_k5(problems, mode=0):
    from rascal_b200 import engine

    return engine.MatchBatch(problems).run(mode=mode).fetch()


def check_against_oracle(r, mx, my, rs, o, deg, mode):
    m = len(o["matched_peaks"])
    assert r.n_matched == m and r.n_ambiguous == o["n_ambiguous"]
    assert np.array_equal(mx[:m], o["matched_peaks"]) and np.array_equal(my[:m], o["matched_atlas"])
    if m <= deg:
        assert not r.refit.accepted
        return
    wl = np.polynomial.polynomial.polyval(o["matched_peaks"], o["polyfit_coeff"])
    got = np.polynomial.polynomial.polyval(o["matched_peaks"], np.array(r.polyfit_coeffs[:deg + 1]))
    assert np.max(np.abs(got - wl)) < 1e-6  # Angstrom
    assert abs(r.polyfit_err - o["polyfit_err"]) < 1e-6 * max(1, m)
    assert r.refit.accepted
    got = np.polynomial.polynomial.polyval(o["matched_peaks"], np.array(r.refit.coeffs[:deg + 1]))
    ref = np.polynomial.polynomial.polyval(o["matched_peaks"], o["fit_coeff"])
    # SciPy's TRF stops 1e-8..1e-5 (relative) short of the optimum, chaotically in its last bits
    # (DESIGN.md section 2): wavelengths to 1e-5 A, the RMS - stationary at the optimum - to 1e-6 A
    assert np.max(np.abs(got - ref)) < 1e-5
    assert abs(r.refit.rms - o["rms"]) < 1e-6
    assert np.max(np.abs(rs[:m] - o["residuals"])) < 1e-5
    if mode == 1:
        exact = O.exact_refit(o["matched_peaks"], o["matched_atlas"], deg)
        if np.max(np.abs(o["matched_atlas"] - np.polynomial.polynomial.polyval(o["matched_peaks"], exact))) < \
                np.std(o["matched_atlas"]):  # all residuals in Huber's quadratic zone
            assert np.allclose(np.array(r.refit.coeffs[:deg + 1]), exact, rtol=1e-9, atol=0)


@pytest.mark.parametrize("case", MATCH_CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_match_peaks_golden(case, mode):
    g = load_match(case)
    deg = len(g["fit_coeff_in"]) - 1
    keys = list(g["status"])
    probs = [dict(peaks=g["peaks"], lines=lines_of(g, k), coeffs=g["fit_coeff_in"], tolerance=tol_of(k), fit_deg=deg)
             for k in keys]
    res, mx, my, rs = run_k5(probs, mode)
    for i, k in enumerate(keys):
        o = O.match_peaks(g["peaks"], lines_of(g, k), g["fit_coeff_in"], tolerance=tol_of(k))
        check_against_oracle(res[i], mx[i], my[i], rs[i], o, deg, mode)
        if g["status"][k] == "ok":  # the reference's own outputs
            m = res[i].n_matched
            assert np.array_equal(mx[i, :m], g["matched_peaks_" + k])
            assert np.array_equal(my[i, :m], g["matched_atlas_" + k])
            if m <= deg:
                continue
            assert abs(res[i].refit.rms - float(g["rms_" + k])) < 1e-6
            assert np.max(np.abs(rs[i, :m] - g["residuals_" + k])) < 1e-5
        else:
            assert res[i].n_ambiguous > 0


def test_match_peaks_ragged_random_batch():
    rng = np.random.default_rng(7)
    probs, oracles = [], []
    for i in range(40):
        n_p, n_a = int(rng.integers(0, 90)), int(rng.integers(0, 400))
        deg = int(rng.integers(1, 6))
        truth = np.array([4000 + 500 * rng.random(), 1 + rng.random(), 2e-5 * rng.standard_normal(),
                          1e-9 * rng.standard_normal(), 1e-13 * rng.standard_normal(),
                          1e-17 * rng.standard_normal()])[:deg + 1]
        peaks = rng.permutation(rng.uniform(0, 4000, n_p))  # caller's order, not sorted
        lines = np.polynomial.polynomial.polyval(rng.uniform(0, 4000, n_a), truth)
        if n_p and n_a > 10 and i % 3 == 0:  # exact hits, duplicated lines and ties
            k = min(4, n_p)
            lines[:k] = np.polynomial.polynomial.polyval(peaks[:k], truth)
            lines[k:2 * k] = lines[:k]
        tol = float(rng.choice([0.5, 2.0, 10.0]))
        probs.append(dict(peaks=peaks, lines=lines, coeffs=truth, tolerance=tol, fit_deg=deg,
                          robust_refit=bool(i % 4)))
        oracles.append(O.match_peaks(peaks, lines, truth, tolerance=tol, fit_deg=deg, robust_refit=bool(i % 4)))
    res, mx, my, rs = run_k5(probs)
    for i, o in enumerate(oracles):
        m = len(o["matched_peaks"])
        assert res[i].n_matched == m and res[i].n_ambiguous == o["n_ambiguous"], i
        assert np.array_equal(mx[i, :m], o["matched_peaks"]) and np.array_equal(my[i, :m], o["matched_atlas"])
        if m > probs[i]["fit_deg"]:
            assert np.max(np.abs(rs[i, :m] - o["residuals"])) < 1e-5, i


def test_match_peaks_device_chained_coeffs():
    """d_coeffs: the coefficients are read from device memory (K4's refit record)."""
    import torch

    g = load_match("poly_quadratic_deg2")
    c = torch.from_numpy(np.ascontiguousarray(g["fit_coeff_in"])).cuda()
    a = run_k5([dict(peaks=g["peaks"], lines=g["lines"], d_coeffs=c.data_ptr(), n_coef=3, fit_deg=2, tolerance=5.0)])
    b = run_k5([dict(peaks=g["peaks"], lines=g["lines"], coeffs=g["fit_coeff_in"], fit_deg=2, tolerance=5.0)])
    assert bytes(a[0]) == bytes(b[0]) and np.array_equal(a[3], b[3])


@pytest.mark.parametrize("case", MATCH_CASES)
def test_facade_match_peaks(case):
    from rascal_b200.calibrator import Calibrator, LineList

    g = load_golden(case)
    gm = load_match(case)
    c = Calibrator(g["peaks"])
    c.set_calibrator_properties(pixel_list=g["pixel_list"])
    key = [k for k, s in gm["status"].items() if s == "ok" and len(gm["matched_peaks_" + k]) > len(g["fit_coeff"])][0]
    c.set_atlas(LineList(lines_of(gm, key)))
    out = c.match_peaks(fit_coeff=g["fit_coeff"], tolerance=tol_of(key))
    assert len(out) == 7
    assert np.array_equal(out[1], gm["matched_peaks_" + key]) and np.array_equal(out[2], gm["matched_atlas_" + key])
    assert abs(out[3] - float(gm["rms_" + key])) < 1e-6
    assert np.allclose(out[5:], gm["utilisation_" + key])
    # inside the data range (a degree-4/5 solution extrapolated to the detector ends amplifies the
    # 1e-8 relative coefficient noise of the TRF stopping point far beyond it)
    wl = np.polynomial.polynomial.polyval(out[1], out[0])
    ref = np.polynomial.polynomial.polyval(out[1], gm["fit_coeff_" + key])
    assert np.max(np.abs(wl - ref)) < 1e-5
    # refine=True (the reference's experimental Nelder-Mead pre-step): with robust_refit its result never reaches
    # the output (calibrator.py:1831 associates with the unrefined coefficients), so the 7-tuple is the same
    out_r = c.match_peaks(fit_coeff=g["fit_coeff"], refine=True, tolerance=tol_of(key))
    assert np.array_equal(out_r[1], out[1]) and np.array_equal(out_r[2], out[2]) and np.allclose(out_r[0], out[0], rtol=1e-12)
    assert c.refine_evaluations > 0
