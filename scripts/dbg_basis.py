import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from tests.conftest import load_golden
from tests.test_gpu_calibrator import make_calibrator
g = load_golden("legendre_deg2")
for mt in (150, 2000, 20000):
    c = make_calibrator(g)
    f = dict(g["config"]["fit"]); f["max_tries"] = mt
    try:
        r = c.fit(progress=False, **f); print(mt, "fit ok", r[0], r[3], len(r[1]))
    except AssertionError as e:
        print(mt, "assert", e)
    print("   ", c.ransac_counters)
