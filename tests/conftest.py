import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["c3_deg4", "poly_linear_deg1", "poly_quadratic_deg2", "fine_deg5", "c1_sprat_xe", "c5_deimos", "c2_gmos_cuar",
                "c4_npix1024", "c4_npix2048", "c4_npix4096", "synth_nm_deg4", "c5_deimos_full", "legendre_deg2",
                "chebyshev_deg2"]


MATCH_CASES = GOLDEN_CASES[:7]  # cases with a match_<case>.npz recording of Calibrator.match_peaks


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    import json

    import numpy as np

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["config"] = json.loads(str(g["config_json"]))
    return g


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)
