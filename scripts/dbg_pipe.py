import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import c4_input
from rascal_b200 import batch
for n, seed in ((12, 77), (12, 5), (20, 77), (11, 77), (13, 77)):
    try:
        r = batch.calibrate_batch([c4_input(i) for i in range(n)], n_hypotheses=20000, seed=seed)
        print(n, seed, "ok", type(r).__name__)
    except Exception as e:
        print(n, seed, "FAIL", e)
