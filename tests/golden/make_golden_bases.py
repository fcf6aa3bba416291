#!/usr/bin/env python
"""Golden fixtures for fit_type='legendre' / 'chebyshev' (calibrator.py:1625-1631), recorded from the UNMODIFIED
reference with the recipe of its own known-answer tests (test/test_polynomial_fit.py:273-364: quadratic atlas,
sample_size = 10, degree 2), at a smaller try budget.

    python tests/golden/make_golden_bases.py     # build container only (needs /root/reference)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the shims, imports the reference)


def main():
    pk, wl, wq = G.polyfit_test_inputs()
    hough = dict(num_slopes=500, range_tolerance=200.0, xbins=100, ybins=100, min_wavelength=3000.0,
                 max_wavelength=8000.0)
    for name, ft in (("legendre_deg2", "legendre"), ("chebyshev_deg2", "chebyshev")):
        G.record_case(name, pk, wq, calib=dict(num_pix=1000), hough=hough,
                      ransac=dict(sample_size=10, minimum_matches=20, minimum_fit_error=1e-12),
                      fit=dict(max_tries=150, fit_tolerance=5.0, candidate_tolerance=2.0, fit_deg=2, fit_type=ft), seed=0)


if __name__ == "__main__":
    main()
