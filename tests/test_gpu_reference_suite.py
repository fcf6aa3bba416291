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


def synthetic_linear(n):
    """SyntheticSpectrum(coefficients=[100, 2]).get_pixels(np.linspace(200, 1200, n)) (synthetic.py:94-124):
    pixels of the wavelengths on the 100 nm + 2 nm/px dispersion, those inside the default 200-1200 range."""
    waves = np.linspace(200, 1200, num=n)
    return (waves - 100.0) / 2.0, waves


def test_synthetic_default():
    from rascal_b200.calibrator import Calibrator, LineList

    peaks, waves = synthetic_linear(20)
    c = Calibrator(peaks=peaks)
    c.set_calibrator_properties(num_pix=768, seed=0)
    c.set_hough_properties(range_tolerance=100.0, min_wavelength=100.0, max_wavelength=1500.0)
    c.set_ransac_properties(minimum_fit_error=1e-12)
    c.set_atlas(LineList(waves))
    best_p = c.fit(max_tries=500, progress=False)[0]
    best_p, x_fit, y_fit, rms, residual, pu, au = c.match_peaks(best_p, refine=False, robust_refit=True)
    fit_diff = c.polyval(x_fit, best_p) - y_fit
    assert pu > 0.7 and au > 0.0 and np.sqrt(np.sum(fit_diff ** 2 / len(x_fit))) < 5.0
