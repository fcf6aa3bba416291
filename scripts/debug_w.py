import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import rascal_oracle as O
from tests.conftest import load_golden
from tests.test_gpu_ransac_refit import setup_from_golden
from rascal_b200 import engine
g = load_golden("poly_quadratic_deg2"); s = setup_from_golden(g)
rb = engine.RansacBatch([s])
n = len(g["tr_status"])
out = rb.run(0, n, replay=(g["tr_x_idx"], g["tr_y_idx"]), per_hypothesis=True)
sc = np.flatnonzero(g["tr_status"] == 4)
d = out["weight"][sc] - g["tr_weight"][sc]
P = np.polynomial.polynomial
px = np.arange(1077.0)
for k in np.flatnonzero(np.abs(d) > 1e-11):
    i = sc[k]
    for nm, c in (("ref", g["tr_coeffs"][i]), ("dev", out["coeffs"][i])):
        t = P.polyval(px, c) - P.polyval(px, O.deriv_coeffs(c)) * px
        cls = np.where(t >= s.gc_max, 2, np.where(t <= s.gc_min, 0, 1))
        print(i, nm, "w", out["weight"][i], g["tr_weight"][i], "cls", np.bincount(cls, minlength=3), "mid t", t[cls == 1], c)
