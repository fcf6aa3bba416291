"""Stage timings of the many-spectra path (BASELINE configs[3] shape)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import c4_input
from rascal_b200 import batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
P = np.polynomial.polynomial
t0 = time.perf_counter(); specs = [c4_input(i) for i in range(n)]; tg = time.perf_counter() - t0
batch.calibrate_batch(specs[:32], n_hypotheses=10_000, seed=1)  # warm
for rep in range(2):
    tm = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = batch.calibrate_batch(specs, n_hypotheses=10_000, seed=1, timings=tm)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    good = fit = 0
    for sp, r in zip(specs, res):
        if r.fit_coeff is None: continue
        fit += 1
        x = np.linspace(0, sp["num_pix"] - 1, 256)
        good += np.sqrt(np.mean((P.polyval(x, r.fit_coeff) - P.polyval(x, sp["truth"])) ** 2)) < 1.0
    print(json.dumps(dict(n=n, gen_s=tg, total_s=dt, spectra_per_s=n / dt, stages=tm, fitted=fit, within_1A=int(good))))
