"""oracle.hough_points_brute_force / hough_histogram against recordings of the reference's
HoughTransform.generate_hough_points_brute_force + bin_hough_points (tests/golden/brute_*.npz)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import rascal_oracle as O
from tests.conftest import GOLDEN_DIR

BRUTE_CASES = ["c3_deg4", "poly_quadratic_deg2"]


def load_brute(case):
    z = np.load(os.path.join(GOLDEN_DIR, "brute_" + case + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def canon(p):
    p = np.asarray(p, dtype=np.float64).reshape(-1, 2)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def sha(p):
    return hashlib.sha256(np.ascontiguousarray(canon(p)).tobytes()).hexdigest()


@pytest.mark.parametrize("case", BRUTE_CASES)
def test_oracle_brute_force_reproduces_reference(case):
    g = load_brute(case)
    pts = O.hough_points_brute_force(g["pairs"][:, 0], g["pairs"][:, 1], *g["limits"])
    assert len(pts) == int(g["n_points"]) and sha(pts) == str(g["points_sorted_sha256"])
    hist, xe, ye = O.hough_histogram(pts, 64, 48)
    assert np.array_equal(hist, g["hist"]) and np.array_equal(xe, g["xedges"]) and np.array_equal(ye, g["yedges"])
    hist, xe, ye = O.hough_histogram(np.vstack((pts, g["extra"])), 40, 40)
    assert np.array_equal(hist, g["hist_added"]) and np.array_equal(xe, g["xedges_added"])
