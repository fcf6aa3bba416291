"""cProfile of the many-spectra path's host side."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import c4_input
from rascal_b200 import batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
specs = [c4_input(i) for i in range(n)]
batch.calibrate_batch(specs[:16], n_hypotheses=10_000, seed=1, match_tolerance=5.0)
tm = {}
batch.calibrate_batch(specs, n_hypotheses=10_000, seed=1, match_tolerance=5.0, timings=tm)
print(tm)
pr = cProfile.Profile(); pr.enable()
batch.calibrate_batch(specs, n_hypotheses=10_000, seed=1, match_tolerance=5.0, timings=tm)
torch.cuda.synchronize()
pr.disable()
print(tm)
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
