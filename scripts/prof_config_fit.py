"""cProfile of the unchanged public call sequence on a BASELINE config: python scripts/prof_config_fit.py c2|c5"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import config_input
from rascal_b200.calibrator import Calibrator, LineList

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = config_input(name)


def run(seed):
    c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
    c.set_calibrator_properties(seed=seed)
    c.set_hough_properties(**w["hough"])
    c.set_ransac_properties(**w["ransac"])
    c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
    c.do_hough_transform()
    r = c.fit(max_tries=w["tries"], progress=False, **w["fit"])
    torch.cuda.synchronize()
    return c


run(1)
t0 = time.perf_counter(); c = run(2); print("one calibration: %.1f ms" % (1e3 * (time.perf_counter() - t0)), c.ransac_counters)
os.environ["CUDA_LAUNCH_BLOCKING"] = "0"
pr = cProfile.Profile(); pr.enable()
run(3)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
