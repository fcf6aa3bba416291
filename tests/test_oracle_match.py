"""oracle.match_peaks against recordings of the reference's Calibrator.match_peaks
(tests/golden/match_*.npz, made by tests/golden/make_golden_match.py)."""
import json
import os

import numpy as np
import pytest

from oracle import rascal_oracle as O
from tests.conftest import GOLDEN_DIR, MATCH_CASES


def load_match(case):
    z = np.load(os.path.join(GOLDEN_DIR, "match_" + case + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["status"] = json.loads(str(g["status_json"]))
    return g


def tol_of(key):
    return float(key.replace("_sep", "").replace("p", "."))


def lines_of(g, key):
    return g["lines_sep"] if key.endswith("_sep") else g["lines"]


@pytest.mark.parametrize("case", MATCH_CASES)
def test_oracle_match_peaks_reproduces_reference(case):
    g = load_match(case)
    n_ok = 0
    for key, st in g["status"].items():
        o = O.match_peaks(g["peaks"], lines_of(g, key), g["fit_coeff_in"], tolerance=tol_of(key))
        if st != "ok":
            # the reference raises exactly when some peak has two lines inside the tolerance
            assert o["n_ambiguous"] > 0
            continue
        n_ok += 1
        assert o["n_ambiguous"] == 0
        assert np.array_equal(o["matched_peaks"], g["matched_peaks_" + key])
        assert np.array_equal(o["matched_atlas"], g["matched_atlas_" + key])
        if len(o["matched_peaks"]) <= len(g["fit_coeff_in"]) - 1:
            continue  # under-determined: the reference returns an arbitrary interpolant; not restated
        # same SciPy call on the same points: bit-identical in one environment
        assert np.array_equal(o["fit_coeff"], g["fit_coeff_" + key])
        assert np.array_equal(o["residuals"], g["residuals_" + key])
        assert o["rms"] == float(g["rms_" + key])
    assert n_ok >= 1
