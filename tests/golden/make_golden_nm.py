#!/usr/bin/env python
"""Golden fixture for the reference's own nm-scale known-answer test (test/test_synthetic_calibration.py:8-62:
20 lines on a 100 nm + 2 nm/px linear dispersion, default degree-4 fit): gradients (~2) and intercepts (~100)
are of the same size as the Hough bin centres, so the Hough-density weight leaves the corner-clamped regime of
SURVEY.md App. A.5 and is a genuine bicubic-spline sum over the pixels (it even goes negative).

    python tests/golden/make_golden_nm.py     # build container only (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the shims, imports the reference)
from rascal.synthetic import SyntheticSpectrum  # noqa: E402


def main():
    s = SyntheticSpectrum(coefficients=[100, 2.0])
    peaks, waves = s.get_pixels(np.linspace(200, 1200, num=20))
    G.record_case("synth_nm_deg4", np.asarray(peaks, float), np.asarray(waves, float), calib=dict(num_pix=768),
                  hough=dict(num_slopes=2000, xbins=100, ybins=100, range_tolerance=100.0, min_wavelength=100.0,
                             max_wavelength=1500.0),  # the three sizes are the defaults, spelled out
                  ransac=dict(minimum_fit_error=1e-12), fit=dict(max_tries=120), seed=0)


if __name__ == "__main__":
    main()
