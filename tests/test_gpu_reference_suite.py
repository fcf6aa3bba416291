"""The reference's own known-answer tests, run through the drop-in facade on the device with the reference's
recipes and pass criteria: test/test_polynomial_fit.py (linear :21-65, quadratic :223-271, legendre :273-318,
chebyshev :320-364) and test/test_synthetic_calibration.py (default :8-62, candidate points from a prior
polynomial + refine :65-121)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def polyfit_inputs():
    """test/test_polynomial_fit.py:5-18."""
    rs = np.random.RandomState(0)
    peaks = np.sort(rs.random_sample(31) * 1000.0)
    m = np.isclose(peaks[:-1], peaks[1:], atol=5.0)
    m = np.insert(m, 0, False)
    peaks = peaks[~m]
    return peaks, 3000.0 + 5.0 * peaks, 3000.0 + 4 * peaks + 1.0e-3 * peaks ** 2.0


def calibrator(peaks, lines, hough, ransac, num_pix=1000, seed=0):
    from rascal_b200.calibrator import Calibrator, LineList

    c = Calibrator(peaks)
    c.set_calibrator_properties(num_pix=num_pix, seed=seed)
    c.set_hough_properties(**hough)
    c.set_atlas(LineList(lines))
    c.set_ransac_properties(**ransac)
    c.do_hough_transform(brute_force=False)
    return c


HOUGH_A = dict(num_slopes=1000, range_tolerance=500.0, xbins=200, ybins=200, min_wavelength=3000.0, max_wavelength=8000.0)
HOUGH_B = dict(num_slopes=500, range_tolerance=200.0, xbins=100, ybins=100, min_wavelength=3000.0, max_wavelength=8000.0)


def test_linear_fit():
    peaks, wl, _ = polyfit_inputs()
    c = calibrator(peaks, wl, HOUGH_A, dict(minimum_matches=20, minimum_fit_error=1e-12))
    best_p = c.fit(max_tries=500, fit_deg=1, progress=False)[0]
    best_p, mp, ma, rms, residual, pu, au = c.match_peaks(best_p, refine=False, robust_refit=True)
    assert np.abs(best_p[1] - 5.0) / 5.0 < 0.001 and np.abs(best_p[0] - 3000.0) / 3000.0 < 0.001
    assert pu > 0.8 and au > 0.0


def test_quadratic_fit():
    peaks, _, wq = polyfit_inputs()
    c = calibrator(peaks, wq, HOUGH_A, dict(minimum_matches=20, minimum_fit_error=1e-12))
    best_p = c.fit(max_tries=2000, fit_tolerance=5.0, candidate_tolerance=2.0, fit_deg=2, progress=False)[0]
    best_p, mp, ma, rms, residual, pu, au = c.match_peaks(best_p, refine=False, robust_refit=True)
    assert np.abs(best_p[2] - 1e-3) / 1e-3 < 0.001 and np.abs(best_p[1] - 4.0) / 4.0 < 0.001
    assert np.abs(best_p[0] - 3000.0) / 3000.0 < 0.001
    assert pu > 0.8 and au > 0.5


@pytest.mark.parametrize("fit_type", ["legendre", "chebyshev"])
def test_quadratic_fit_other_bases(fit_type):
    """legfit / chebfit on the raw pixels inside the loop, the monomial robust refit as the result - which is why
    the reference's asserts read the returned coefficients as monomials (:313-318)."""
    peaks, _, wq = polyfit_inputs()
    c = calibrator(peaks, wq, HOUGH_B, dict(sample_size=10, minimum_matches=20, minimum_fit_error=1e-12))
    best_p, mp, ma, rms, residual, pu, au = c.fit(max_tries=2000, fit_tolerance=5.0, candidate_tolerance=2.0, fit_deg=2,
                                                  fit_type=fit_type, progress=False)
    assert np.abs(best_p[2] - 1e-3) / 1e-3 < 0.001 and np.abs(best_p[1] - 4.0) / 4.0 < 0.001
    assert np.abs(best_p[0] - 3000.0) / 3000.0 < 0.001
    assert pu > 0.7 and au > 0.5
    assert c.ransac_counters["unsupported"] == 0


def synthetic_linear():
    """SyntheticSpectrum(coefficients=[100, 2]).get_pixels(np.linspace(200, 1200, 20)) as the reference evaluated it
    (synthetic.py:94-124 keeps the wavelengths strictly inside its 200-1200 nm range: 18 of the 20), taken from
    the recording of that run."""
    from tests.conftest import load_golden

    g = load_golden("synth_nm_deg4")
    return g["peaks"], g["lines"]


def test_synthetic_default():
    from rascal_b200.calibrator import Calibrator, LineList

    peaks, waves = synthetic_linear()
    c = Calibrator(peaks=peaks)
    c.set_calibrator_properties(num_pix=768, seed=0)
    c.set_hough_properties(range_tolerance=100.0, min_wavelength=100.0, max_wavelength=1500.0)
    c.set_ransac_properties(minimum_fit_error=1e-12)
    c.set_atlas(LineList(waves))
    best_p = c.fit(max_tries=500, progress=False)[0]
    best_p, x_fit, y_fit, rms, residual, pu, au = c.match_peaks(best_p, refine=False, robust_refit=True)
    fit_diff = c.polyval(x_fit, best_p) - y_fit
    assert pu > 0.7 and au > 0.0 and np.sqrt(np.sum(fit_diff ** 2 / len(x_fit))) < 5.0


def test_get_candidate_points_poly_and_refine():
    """test/test_synthetic_calibration.py:65-121: set_ransac_properties(linear=False) - candidates from the prior
    polynomial (calibrator.py:263-315) - then match_peaks(refine=True) (:1797-1821).  Same seed as the recording
    of that recipe: the candidate table is the reference's, and so are its pass criteria."""
    from rascal_b200.calibrator import Calibrator, LineList
    from tests.conftest import load_golden

    g = load_golden("poly_candidates_nm")
    best_p = [100, 2.0]
    c = Calibrator(peaks=g["peaks"])
    c.set_calibrator_properties(num_pix=768, seed=0)
    c.set_hough_properties(range_tolerance=100.0, min_wavelength=100.0, max_wavelength=1500.0)
    c.set_ransac_properties(linear=False, minimum_fit_error=1e-12)
    c.set_atlas(LineList(g["lines"]))
    best_p, x, y, rms, residual, pu, au = c.fit(max_tries=500, fit_coeff=best_p, progress=False)
    assert np.array_equal(np.array(c.candidate_peak), g["candidate_peak"])
    assert np.array_equal(np.array(c.candidate_arc), g["candidate_arc"])
    assert np.array_equal([len(t[0]) for t in c.candidates], g["n_cand_per_line"])
    best_p, x_fit, y_fit, rms, residual, pu, au = c.match_peaks(best_p, refine=True, robust_refit=True)
    fit_diff = c.polyval(x_fit, best_p) - y_fit
    assert pu > 0.7 and au > 0.0 and np.sqrt(np.sum(fit_diff ** 2 / len(x_fit))) < 5.0
    assert c.refine_evaluations > 10
    # robust_refit=False returns the Nelder-Mead result itself (:1944)
    out = c.match_peaks(np.array([100.0, 2.0]), refine=True, robust_refit=False)
    assert np.allclose(out[0], g["refine_coeff_out"], rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError, match="guess solution"):
        c2 = Calibrator(peaks=g["peaks"])
        c2.set_calibrator_properties(num_pix=768, seed=0)
        c2.set_hough_properties(range_tolerance=100.0, min_wavelength=100.0, max_wavelength=1500.0)
        c2.set_ransac_properties(linear=False)
        c2.set_atlas(LineList(g["lines"]))
        c2.fit(max_tries=50, progress=False)


def test_adjust_polyfit_objective_equals_reference():
    """rb2_adjust_polyfit against every evaluation the reference's Nelder-Mead made and against probes off its path
    (incl. the +inf branches: too few matches, non-monotonic)."""
    from rascal_b200 import engine
    from tests.conftest import load_golden

    g = load_golden("poly_candidates_nm")
    obj = engine.AdjustPolyfit(g["peaks"], g["lines"], g["pixel_list"], g["refine_fit_in"], float(g["refine_tolerance"]),
                               float(g["refine_min_frac"]))
    for deltas, ref in ((g["refine_delta"], g["refine_lsq"]), (g["probe_delta2"], g["probe_lsq2"]),
                        (g["probe_delta1"], g["probe_lsq1"])):
        got = obj.many(deltas)
        assert np.array_equal(np.isinf(got), np.isinf(ref))
        fin = np.isfinite(ref)
        assert np.allclose(got[fin], ref[fin], rtol=1e-12, atol=1e-300)
    assert obj(g["refine_delta"][0]) == obj.many(g["refine_delta"][:1])[0]
