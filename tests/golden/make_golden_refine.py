#!/usr/bin/env python
"""Golden fixture for the two experimental options of the reference, recorded from the
UNMODIFIED reference with the recipe of its own test (test/test_synthetic_calibration.py:65-121):

  * set_ransac_properties(linear=False): candidates from a prior polynomial, Calibrator._get_candidate_points_poly
    (calibrator.py:263-315) - the candidate table and the seeded RANSAC trace;
  * match_peaks(refine=True): scipy.optimize.minimize(Nelder-Mead) over Calibrator._adjust_polyfit
    (calibrator.py:754-820, 1797-1821) - every (delta, lsq) evaluation and the result.

    python tests/golden/make_golden_refine.py     # build container only (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the shims, imports the reference)
from rascal.atlas import Atlas  # noqa: E402
from rascal.calibrator import Calibrator  # noqa: E402
from rascal.synthetic import SyntheticSpectrum  # noqa: E402


def main():
    best_p = [100, 2.0]
    s = SyntheticSpectrum(coefficients=best_p)
    peaks, waves = s.get_pixels(np.linspace(200, 1200, num=25))
    hough = dict(num_slopes=2000, xbins=100, ybins=100, range_tolerance=100.0, min_wavelength=100.0,
                 max_wavelength=1500.0)
    G.record_case("poly_candidates_nm", np.asarray(peaks, float), np.asarray(waves, float), calib=dict(num_pix=768),
                  hough=hough, ransac=dict(linear=False, minimum_fit_error=1e-12),
                  fit=dict(max_tries=100, fit_coeff=best_p), seed=0)
    # the refine step of the same test, on a fresh seeded calibrator
    c = Calibrator(peaks)
    c.set_calibrator_properties(num_pix=768, seed=0)
    c.set_hough_properties(**hough)
    c.set_ransac_properties(linear=False, minimum_fit_error=1e-12)
    a = Atlas()
    a.add_user_atlas(elements=["Test"] * len(waves), wavelengths=waves)
    c.set_atlas(a)
    fit = c.fit(max_tries=100, fit_coeff=best_p, progress=False)[0]
    calls = []
    orig = Calibrator._adjust_polyfit

    def logged(self, delta, fit_, tolerance, min_frac):
        v = orig(self, delta, fit_, tolerance, min_frac)
        calls.append((np.array(delta, float), float(v)))
        return v

    Calibrator._adjust_polyfit = logged
    try:
        out = c.match_peaks(fit, refine=True, robust_refit=False)
    finally:
        Calibrator._adjust_polyfit = orig
    # probes of the objective away from the optimiser's path: shifted intercepts / slopes, incl. infeasible ones
    rng = np.random.default_rng(1)
    probes = [np.array([d0, d1]) for d0, d1 in zip(rng.normal(0, 30, 40), rng.normal(0, 0.5, 40))]
    probes += [np.array([d0]) for d0 in rng.normal(0, 3, 10)]
    probe_val = [float(orig(c, p, np.asarray(fit, float), 10.0, 0.5)) for p in probes]
    z = dict(np.load(os.path.join(HERE, "poly_candidates_nm.npz"), allow_pickle=False))
    z.update(refine_fit_in=np.asarray(fit, float), refine_delta=np.array([d for d, _ in calls]),
             refine_lsq=np.array([v for _, v in calls]), refine_coeff_out=np.asarray(out[0], float),
             refine_tolerance=np.float64(10.0), refine_min_frac=np.float64(0.5),
             probe_delta2=np.array(probes[:40]), probe_lsq2=np.array(probe_val[:40]),
             probe_delta1=np.array(probes[40:]), probe_lsq1=np.array(probe_val[40:]))
    np.savez_compressed(os.path.join(HERE, "poly_candidates_nm.npz"), **z)
    print("refine: %d objective evaluations, coefficients %s -> %s" % (len(calls), fit, out[0]))


if __name__ == "__main__":
    main()
