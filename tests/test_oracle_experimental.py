"""Oracle restatements of the reference's two experimental options against the recording
of its own test recipe (tests/golden/poly_candidates_nm.npz, make_golden_refine.py):
linear=False candidates (calibrator.py:263-315) and the refine objective _adjust_polyfit (:754-820)."""
import numpy as np

from oracle import rascal_oracle as O
from tests.conftest import load_golden


def test_candidate_points_poly_reproduce_reference():
    g = load_golden("poly_candidates_nm")
    cfg = g["config"]
    rs = np.random.RandomState(cfg["seed"])
    cands = O.candidate_points_poly(g["peaks"], g["lines"], cfg["fit"]["fit_coeff"], g["hist"].size,
                                    cfg["hough"]["range_tolerance"], 2.0, rs)
    assert np.array_equal([len(t[0]) for t in cands], g["n_cand_per_line"])
    cp, ca = O.most_common_candidates(cands, 5, True)
    assert np.array_equal(cp, g["candidate_peak"]) and np.array_equal(ca, g["candidate_arc"])


def test_adjust_polyfit_reproduces_reference():
    g = load_golden("poly_candidates_nm")
    args = (g["refine_fit_in"], g["peaks"], g["lines"], g["pixel_list"], float(g["refine_tolerance"]), float(g["refine_min_frac"]))
    for d, v in zip(g["refine_delta"], g["refine_lsq"]):
        assert O.adjust_polyfit(d, *args) == v
    for d, v in list(zip(g["probe_delta2"], g["probe_lsq2"])) + list(zip(g["probe_delta1"], g["probe_lsq1"])):
        assert O.adjust_polyfit(d, *args) == v
    assert np.isinf(g["probe_lsq2"]).any() and np.isfinite(g["probe_lsq2"]).any()
