import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import c4_input
from rascal_b200 import batch
specs = [c4_input(i) for i in range(64)]
res = batch.calibrate_batch(specs, n_hypotheses=10_000, seed=1)
print("done", sum(r.fit_coeff is not None for r in res))
