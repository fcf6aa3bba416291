"""K4 (rb2_refit_winners) alone on the recorded C3 winner: CUDA-event time of the kernel, and - with a library built
with -DRB2_K4_PROFILE (RB2_LIB=rascal_b200/_variants/librascal_b200_k4prof.so) - its own cycle breakdown."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rascal_b200 import engine
from tests.conftest import load_golden
from tests.test_gpu_ransac_refit import setup_from_golden

for case in (sys.argv[1:] or ["c3_deg4", "c2_gmos_cuar"]):
    g = load_golden(case)
    s = setup_from_golden(g)
    rb = engine.RansacBatch([s])
    n = len(g["tr_status"])
    rb.run(0, n, replay=(g["tr_x_idx"], g["tr_y_idx"]))
    for mode in (0, 1):
        rb.refit_async(mode)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); rb.refit_async(mode); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        fit = rb.fetch()[1][0]
        print("%s mode %d: refit_kernel %.1f us (median of 5), m = %d inliers, deg %d, nfev %d, status %d"
              % (case, mode, sorted(ts)[2], fit.n_inliers, s.deg, fit.huber_iters, fit.pad))
