"""Find hypotheses whose device cost disagrees with the oracle (debug)."""
import json, sys, os
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
rb = engine.RansacBatch([s])
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
out = rb.run(0, n, seed=99, per_hypothesis=True)
sc = np.flatnonzero(out["status"] == 4)
print("scored", len(sc))
order = sc[np.argsort(out["cost"][sc])]
bad = []
for i in list(order[:300]) + list(sc[:1500]):
    c = out["coeffs"][i]
    err, mx, my = O.match_bijective(prob, c)
    err[err > prob.tol] = prob.tol
    wt = O.hough_weight_sum(prob, c)
    cost = sum(err) / (len(err) - 5 + 1) / (wt + 1e-9)
    if not np.isclose(cost, out["cost"][i], rtol=1e-9) or out["counts"][i, 0] != len(err):
        bad.append(dict(i=int(i), dev_cost=float(out["cost"][i]), cost=float(cost), dev_w=float(out["weight"][i]), w=float(wt),
                        dev_counts=out["counts"][i].tolist(), m=len(err), inl=int((err < prob.tol).sum()), coeffs=c.tolist()))
print("bad", len(bad))
for b in bad[:10]:
    print(b)
json.dump(bad, open("gpurun_out/bad_cost.json", "w"))
