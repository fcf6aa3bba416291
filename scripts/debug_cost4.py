import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess
from rascal_b200.csrc import build as b
b.NVCC_FLAGS.append("-DRB2_DEBUG_WEIGHT")
b.build(force=True)
import numpy as np
from bench import c3_input
from rascal_b200 import engine
from rascal_b200.calibrator import Calibrator, LineList
w = c3_input()
c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
c.set_hough_properties(**w["hough"]); c.set_ransac_properties(**w["ransac"])
c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
c.do_hough_transform(); c._vote(10.0)
setup = engine.RansacSetup(c.candidate_peak, c.candidate_arc, fit_deg=4, sample_size=5, ransac_tolerance=5,
    min_intercept=c.min_intercept, max_intercept=c.max_intercept, pixel_list=c.pixel_list, hist=c.ht.hist,
    xedges=c.ht.xedges, yedges=c.ht.yedges, hough_weight=1.0, n_peaks_total=60)
rb = engine.RansacBatch([setup], blocks_per_problem=1)
out = rb.run(956017, 1, seed=9171440914760070181, per_hypothesis=True)
import torch; torch.cuda.synchronize()
print(out)
