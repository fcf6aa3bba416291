import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import c3_input
from rascal_b200 import engine
from rascal_b200.calibrator import Calibrator, LineList
w = c3_input()
c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
c.set_calibrator_properties(seed=1)
c.set_hough_properties(**w["hough"]); c.set_ransac_properties(**w["ransac"])
c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
c.do_hough_transform()
c._vote(10.0)
np.random.seed(1)
key = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
print("key", key)
setup = engine.RansacSetup(c.candidate_peak, c.candidate_arc, fit_deg=4, sample_size=5, ransac_tolerance=5,
    min_intercept=c.min_intercept, max_intercept=c.max_intercept, pixel_list=c.pixel_list, hist=c.ht.hist,
    xedges=c.ht.xedges, yedges=c.ht.yedges, hough_weight=1.0, n_peaks_total=60)
rb = engine.RansacBatch([setup])
n = 1000000
out = rb.run(0, n, seed=key, per_hypothesis=True)
w1 = rb.results()[0]
print("record", w1.hyp_id, w1.cost, w1.n_match, w1.n_inliers, list(w1.coeffs[:5]))
k = w1.hyp_id
print("per-hyp at id", out["status"][k], out["cost"][k], out["weight"][k], out["counts"][k], out["coeffs"][k])
sc = np.flatnonzero(out["status"] == 4)
o = sc[np.argsort(out["cost"][sc])][:8]
for i in o: print(i, out["cost"][i], out["weight"][i], out["counts"][i], out["coeffs"][i])
