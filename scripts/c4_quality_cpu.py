"""Oracle (reference algorithm, CPU) on the first C4 spectra: success vs truth at ~equal draw budget."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import numpy as np
from bench import c4_input
from oracle import rascal_oracle as O
P = np.polynomial.polynomial
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tries = int(sys.argv[2]) if len(sys.argv) > 2 else 250
good = fit = 0
for i in range(n):
    sp = c4_input(i)
    trace = []
    t0 = time.perf_counter()
    try:
        r = O.run_path(sp["peaks"], sp["lines"], num_pix=sp["num_pix"], max_tries=tries, seed=i, trace=trace,
                       **sp["hough"], **sp["ransac"], **sp["fit"])
    except Exception as e:
        print(i, "failed", type(e).__name__, e); continue
    b = r["best"]
    c = b["best_p"]
    if c is None:
        print(i, "no fit, draws", len(trace)); continue
    fit += 1
    x = np.linspace(0, sp["num_pix"] - 1, 256)
    e = np.sqrt(np.mean((P.polyval(x, c) - P.polyval(x, sp["truth"])) ** 2))
    good += e < 1.0
    print(i, "draws %d err %.2f A  %.1f s" % (len(trace), e, time.perf_counter() - t0), flush=True)
print("fitted %d/%d within 1 A: %d" % (fit, n, good))
