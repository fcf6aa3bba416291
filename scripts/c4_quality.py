"""Success rate of the many-spectra path on C4 spectra vs hypothesis budget."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import c4_input
from rascal_b200 import batch
P = np.polynomial.polynomial
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
specs = [c4_input(i) for i in range(n)]
for H in (10_000, 100_000, 1_000_000):
    t0 = time.perf_counter()
    res = batch.calibrate_batch(specs, n_hypotheses=H, seed=1, match_tolerance=5.0)
    dt = time.perf_counter() - t0
    good = good_m = fit = 0
    rms = []
    for sp, r in zip(specs, res):
        if r.fit_coeff is None:
            continue
        fit += 1
        x = np.linspace(0, sp["num_pix"] - 1, 256)
        e = np.sqrt(np.mean((P.polyval(x, r.fit_coeff) - P.polyval(x, sp["truth"])) ** 2))
        rms.append(e)
        good += e < 1.0
        if r.match is not None:
            good_m += np.sqrt(np.mean((P.polyval(x, r.match.fit_coeff) - P.polyval(x, sp["truth"])) ** 2)) < 1.0
    print("H=%d: fitted %d/%d, within 1 A: %d (after match_peaks: %d), median err %.2f A, %.2f s" %
          (H, fit, n, good, good_m, float(np.median(rms)), dt))
