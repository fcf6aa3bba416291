"""cProfile of the single-spectrum public path (do_hough_transform + fit) on the C3 workload."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import c3_input
from rascal_b200.calibrator import Calibrator, LineList

w = c3_input()
H = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000


def new(seed):
    c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
    c.set_calibrator_properties(seed=seed)
    c.set_hough_properties(**w["hough"])
    c.set_ransac_properties(**w["ransac"])
    c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
    return c


def run(seed):
    c = new(seed)
    c.do_hough_transform()
    r = c.fit(progress=False, n_hypotheses=H, **w["fit"])
    torch.cuda.synchronize()
    return r


run(1); run(2)
for k in range(3):
    t0 = time.perf_counter(); c = new(10 + k); t1 = time.perf_counter()
    c.do_hough_transform(); torch.cuda.synchronize(); t2 = time.perf_counter()
    c.fit(progress=False, n_hypotheses=H, **w["fit"]); torch.cuda.synchronize(); t3 = time.perf_counter()
    print("construct %.2f ms  hough %.2f ms  fit %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
pr = cProfile.Profile(); pr.enable()
for k in range(5):
    run(20 + k)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
