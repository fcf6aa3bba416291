"""Random small calibration problems for differential testing against the reference (tests/test_oracle_live_reference.py
draws new ones and lets the reference solve them on the spot; tests/golden/make_golden_live.py recorded the
LIVE_CASES below as fixtures so that the device path is held to them on the GPU box, where the reference is absent)."""
import numpy as np

# seeds of the committed fixtures tests/golden/live_<seed>.npz - chosen to cover what the other goldens are thin on:
# candidate_weighted=False, filter_close, over-determined samples, hough_weight != 1 and the other bases at degree 3 / 4
LIVE_CASES = [20262, 20264, 30002, 30003, 30006, 30008, 30010, 31003, 31007]
# the same with set_known_pairs (random_problem(seed, with_known=True)): fixtures tests/golden/live_known_<seed>.npz
KNOWN_CASES = [32000, 32001, 32005]


def random_problem(seed, with_known=False):
    """A small synthetic arc in the style of SURVEY.md 8d (true lines + distractors + spurious peaks), every
    setting drawn from `seed`."""
    rng = np.random.default_rng(seed)
    npix = int(rng.choice([600, 1024, 1500]))
    deg = int(rng.choice([2, 3, 4]))
    co = [float(rng.uniform(3500, 7000)), float(rng.uniform(0.8, 2.5)), float(rng.uniform(-4e-5, 4e-5)),
          float(rng.uniform(-6e-9, 6e-9)), 0.0]
    lam = lambda x: np.polynomial.polynomial.polyval(x, co)  # noqa: E731
    n_peaks = int(rng.integers(14, 26))
    px = np.sort(rng.uniform(10, npix - 10, n_peaks))
    px = px[np.concatenate(([True], np.diff(px) > 3.0))]
    peaks = px + rng.normal(0, 0.1, len(px))
    keep = rng.random(len(px)) > 0.15  # some peaks have no atlas line
    lines = np.sort(np.concatenate([lam(px[keep]), rng.uniform(lam(0) - 200, lam(npix - 1) + 200, int(rng.integers(10, 40)))]))
    over = int(rng.choice([0, 0, 1, 3]))  # over-determined samples now and then
    spec = dict(
        name="live_%d" % seed, peaks=peaks.tolist(), lines=lines.tolist(), npix=npix, seed=int(rng.integers(0, 1000)),
        hough=dict(num_slopes=int(rng.choice([300, 500, 800])), xbins=int(rng.choice([40, 64, 100])),
                   ybins=int(rng.choice([40, 50, 90])), min_wavelength=float(lam(0) + rng.uniform(-40, 40)),
                   max_wavelength=float(lam(npix - 1) + rng.uniform(-40, 40)),
                   range_tolerance=float(rng.choice([200, 350, 500])), linearity_tolerance=float(rng.choice([50, 100]))),
        ransac=dict(sample_size=deg + 1 + over, top_n_candidate=int(rng.choice([3, 5, 8])),
                    ransac_tolerance=float(rng.choice([2, 5])), filter_close=bool(rng.random() < 0.3),
                    minimum_matches=int(rng.choice([1, 6])), minimum_peak_utilisation=float(rng.choice([0.0, 0.2])),
                    candidate_weighted=bool(rng.random() < 0.8), hough_weight=float(rng.choice([1.0, 0.5]))),
        fit=dict(max_tries=int(rng.integers(25, 60)), fit_deg=deg, candidate_tolerance=float(rng.choice([4.0, 8.0]))))
    fit_type = str(rng.choice(["poly", "poly", "legendre", "chebyshev"]))  # drawn last: earlier draws keep their values
    if fit_type != "poly":
        spec["fit"]["fit_type"] = fit_type
    if with_known and rng.random() < 0.6:
        # set_known_pairs (calibrator.py:1507-1550): a few (pixel, wavelength) pairs the user is sure of, appended to
        # every sample before the fit (:568-573).  (Not part of the recorded LIVE_CASES fixtures.)
        k = int(rng.integers(1, 4))
        kp = rng.uniform(20, npix - 20, k)
        spec["known"] = [kp.tolist(), (lam(kp) + rng.normal(0, 0.05, k)).tolist()]
    return spec
