"""Wall time of the default public call sequence (do_hough_transform + fit(max_tries=...)) on C3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import c3_input
from rascal_b200.calibrator import Calibrator, LineList

w = c3_input()
def new(seed):
    c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
    c.set_calibrator_properties(seed=seed)
    c.set_hough_properties(**w["hough"])
    c.set_ransac_properties(**w["ransac"])
    c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
    return c
for mt in (500, 5000):
    for k in range(4):
        t0 = time.perf_counter(); c = new(k); c.do_hough_transform()
        r = c.fit(max_tries=mt, progress=False, **w["fit"]); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("max_tries %d: %.2f ms per calibration, rms %.3f, counters %s" % (mt, dt * 1e3, r[3], c.ransac_counters))
