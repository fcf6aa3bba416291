"""Host-side time of one resident bench step (rb.run / refit_async / fetch) next to its device time: is a small
step (a rank's share of an 8-GPU run) bound by the CPU's launch path?   usage: python scripts/prof_step_host.py [hyp]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import c3_input
from rascal_b200 import engine
from rascal_b200.calibrator import Calibrator, LineList

H = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12_500_000
w = c3_input()
c = Calibrator(w["peaks"], spectrum=np.zeros(w["npix"]))
c.set_calibrator_properties(seed=1)
c.set_hough_properties(**w["hough"])
c.set_ransac_properties(**w["ransac"])
c.set_atlas(LineList(w["atlas"]), candidate_tolerance=10.0)
c.do_hough_transform()
c._vote(10.0)
setup = engine.RansacSetup(c.candidate_peak, c.candidate_arc, fit_deg=4, sample_size=5, ransac_tolerance=c.ransac_tolerance,
                           min_intercept=c.min_intercept, max_intercept=c.max_intercept, pixel_list=c.pixel_list, hist=c.ht.hist,
                           xedges=c.ht.xedges, yedges=c.ht.yedges, hough_weight=c.hough_weight, n_peaks_total=len(c.peaks))
rb = engine.RansacBatch([setup])
for k in range(3):
    rb.run(k * H, H, seed=2026); rb.refit_async(); rb.fetch()
torch.cuda.synchronize()
rows = []
for k in range(8):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e[0].record()
    rb.run((3 + k) * H, H, seed=2026); t1 = time.perf_counter(); e[1].record()
    rb.refit_async(); t2 = time.perf_counter(); e[2].record()
    rb.fetch(); t3 = time.perf_counter(); e[3].record()
    torch.cuda.synchronize()
    rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])))
r = np.median(np.array(rows), axis=0)
print("hyp %.3g: host ms  run %.3f  refit_async %.3f  fetch(sync) %.3f | device ms  K3 %.3f  K4 %.3f  D2H %.3f | step wall %.3f"
      % (H, r[0], r[1], r[2], r[3], r[4], r[5], r[0] + r[1] + r[2]))
# K4 alone, warm vs after an L2 flush (the bench flushes L2 between steps: the kernel's own code is then cold too)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for cold in (False, True):
    ts = []
    for _ in range(7):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rb.refit_async(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print("K4 %s L2: %.1f us (median of 7)" % ("after a flushed" if cold else "warm", sorted(ts)[3]))
