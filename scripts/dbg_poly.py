import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from tests.conftest import load_golden
from rascal_b200.calibrator import Calibrator, LineList
from rascal_b200 import engine
g = load_golden("poly_candidates_nm")
c = Calibrator(peaks=g["peaks"])
c.set_calibrator_properties(num_pix=768, seed=0)
c.set_hough_properties(range_tolerance=100.0, min_wavelength=100.0, max_wavelength=1500.0)
c.set_ransac_properties(linear=False, minimum_fit_error=1e-12)
c.set_atlas(LineList(g["lines"]))
try:
    print(c.fit(max_tries=500, fit_coeff=[100, 2.0], progress=False)[0])
except AssertionError as e:
    print("assert", e)
print(c.ransac_counters, len(c.candidate_peak))
s = engine.RansacSetup(c.candidate_peak, c.candidate_arc, fit_deg=1, sample_size=5, ransac_tolerance=5, min_intercept=c.min_intercept,
    max_intercept=c.max_intercept, pixel_list=c.pixel_list, hist=c.ht.hist, xedges=c.ht.xedges, yedges=c.ht.yedges, minimum_fit_error=1e-12, n_peaks_total=len(c.peaks))
print("prior cost", s.prior_cost([100, 2.0]), "general", s.bs_c is not None)
rb = engine.RansacBatch([s])
out = rb.run(0, 2000, seed=1, per_hypothesis=True)
print(engine.counters_dict(rb.results()[0]), rb.results()[0].cost, rb.results()[0].hyp_id)
sc = out["status"] == 4
print("scored", sc.sum(), "costs", np.sort(out["cost"][sc])[:5], "weights", out["weight"][sc][:5], "counts", out["counts"][sc][:3])
fit, mx, my, rs = rb.refit()
print("refit", fit[0].accepted, fit[0].rms, fit[0].n_inliers, list(fit[0].coeffs[:2]))
