import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import c3_input
from oracle import rascal_oracle as O
from rascal_b200 import engine
w = c3_input()
r = O.run_path(w["peaks"], w["atlas"], num_pix=w["npix"], max_tries=0, seed=0, **w["hough"], **w["ransac"], **w["fit"])
prob = r["problem"]
s = engine.RansacSetup(r["candidate_peak"], r["candidate_arc"], fit_deg=4, sample_size=5, ransac_tolerance=5,
                       min_intercept=prob.min_intercept, max_intercept=prob.max_intercept, pixel_list=prob.pixel_list,
                       hist=r["hist"], xedges=r["xedges"], yedges=r["yedges"], n_peaks_total=60)
for bpp in (0, 1, 8):
    rb = engine.RansacBatch([s], blocks_per_problem=bpp)
    for n in (10000, 200000, 1000000):
        out = rb.run(0, n, seed=99, per_hypothesis=True)
        w1 = rb.results()[0]
        sc = np.flatnonzero(out["status"] == 4)
        el = sc[out["counts"][sc, 1] >= 5]
        j = el[np.lexsort((-el, out["cost"][el]))[0]]
        rb.run(0, n, seed=99)
        w2 = rb.results()[0]
        print(bpp, n, "per-hyp best", j, out["cost"][j], "| record(with outputs)", w1.hyp_id, w1.cost, w1.n_match, w1.n_inliers,
              "| record(no outputs)", w2.hyp_id, w2.cost, w2.n_match, w2.n_inliers)
        if w2.hyp_id >= 0 and w2.hyp_id != j:
            k = w2.hyp_id
            print("    at record id: status", out["status"][k], "cost", out["cost"][k], "counts", out["counts"][k], "coeffs equal", np.array_equal(out["coeffs"][k], np.array(w2.coeffs[:5])))
