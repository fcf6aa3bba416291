"""Dump the 1e8-hypothesis winner of the C3 problem for offline inspection."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import c3_input
from rascal_b200.calibrator import Calibrator, LineList
w = c3_input()
out = {}
for H in (10**4, 10**6, 10**8):
    c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
    c.set_calibrator_properties(seed=1)
    c.set_hough_properties(**w["hough"]); c.set_ransac_properties(**w["ransac"])
    c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
    r = c.fit(progress=False, n_hypotheses=H, **w["fit"])
    out[str(H)] = dict(coeff=list(r[0]), rms=float(r[3]), n=len(r[1]), mp=list(map(float, r[1])), ma=list(map(float, r[2])),
                       cost=c.best_cost, hyp=[int(c.best_hypothesis[0]), list(c.best_hypothesis[1])], counters=c.ransac_counters,
                       cand_peak=list(map(float, c.candidate_peak)), cand_arc=list(map(float, c.candidate_arc)))
json.dump(out, open("gpurun_out/winner.json", "w"))
print({k: (v["n"], v["rms"], v["cost"]) for k, v in out.items()})
