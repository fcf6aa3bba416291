"""Explicit Hough points on the device: rb2_hough_brute_force and
rb2_hough_bin_points against recordings of the reference and the oracle - bit-exact points
(as a multiset; the reference's row order depends on an unstable argsort), histogram, edges,
canonical ranking; add_hough_points / save / load round trip; brute-force front-end of the facade."""
import numpy as np
import pytest

from oracle import rascal_oracle as O
from tests.conftest import load_golden
from tests.test_oracle_brute import BRUTE_CASES, canon, load_brute, sha

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", BRUTE_CASES)
def test_brute_force_points_and_binning(case):
    from rascal_b200.houghtransform import HoughTransform

    g = load_brute(case)
    ht = HoughTransform()
    ht.set_constraints(*g["limits"])
    ht.generate_hough_points_brute_force(g["pairs"][:, 0], g["pairs"][:, 1])
    pts = np.asarray(ht.hough_points)
    assert len(pts) == int(g["n_points"]) and sha(pts) == str(g["points_sorted_sha256"])
    # row order: the reference's, for a stable sort of x
    order = np.argsort(g["pairs"][:, 0], kind="stable")
    x, y = g["pairs"][order, 0], g["pairs"][order, 1]
    ref = []
    for i in range(len(x) - 1):
        gr = (y[i + 1:] - y[i]) / (x[i + 1:] - x[i])
        ic = y[i + 1:] - gr * x[i + 1:]
        k = (gr > g["limits"][0]) & (gr < g["limits"][1]) & (ic > g["limits"][2]) & (ic < g["limits"][3])
        ref.append(np.column_stack((gr[k], ic[k])))
    assert np.array_equal(pts, np.vstack(ref))
    ht.bin_hough_points(64, 48)
    assert np.array_equal(ht.hist, g["hist"]) and np.array_equal(ht.xedges, g["xedges"])
    assert np.array_equal(ht.yedges, g["yedges"])
    flat = np.ravel_multi_index((ht.hist_sorted_arg[:, 0], ht.hist_sorted_arg[:, 1]), ht.hist.shape)
    assert np.array_equal(flat, O.rank_bins(g["hist"]))
    ht.add_hough_points(g["extra"])
    ht.bin_hough_points(40, 40)
    assert np.array_equal(ht.hist, g["hist_added"]) and np.array_equal(ht.xedges, g["xedges_added"])
    assert np.array_equal(ht.yedges, g["yedges_added"])


def test_bin_points_random_and_edge_cases():
    from rascal_b200 import engine

    rng = np.random.default_rng(11)
    probs = []
    for n, xb, yb in ((0, 5, 7), (1, 1, 1), (1, 4, 3), (5000, 100, 100), (777, 3, 250), (2, 2, 2)):
        pts = np.column_stack((rng.uniform(0.5, 3.0, n), rng.uniform(3000, 9000, n)))
        if n > 100:
            pts[:50] = pts[50:100]          # duplicates
            pts[100, :] = pts[:, :].max(axis=0)  # a point on both last edges
        probs.append(dict(points=pts, xbins=xb, ybins=yb))
    hb = engine.PointsHoughBatch(probs).run()
    for i, p in enumerate(probs):
        hist, xe, ye = np.histogram2d(p["points"][:, 0], p["points"][:, 1], bins=(p["xbins"], p["ybins"]))
        assert np.array_equal(hb.hist(i), hist.astype(np.uint32)), i
        assert np.array_equal(hb.xedges(i), xe) and np.array_equal(hb.yedges(i), ye), i
        assert np.array_equal(hb.rank(i), O.rank_bins(hist)), i
    with pytest.raises(ValueError):
        engine.PointsHoughBatch([dict(points=np.array([[1.0, np.nan]]), xbins=2, ybins=2)])


def test_save_load_round_trip_rebins(tmp_path):
    from rascal_b200.houghtransform import HoughTransform

    g = load_brute("poly_quadratic_deg2")
    a = HoughTransform()
    a.set_constraints(*g["limits"])
    a.generate_hough_points_brute_force(g["pairs"][:, 0], g["pairs"][:, 1])
    a.bin_hough_points(64, 48)
    for fmt in ("npy", "json"):
        a.save(str(tmp_path / "ht"), fileformat=fmt)
        b = HoughTransform()
        b.load(str(tmp_path / "ht"), filetype=fmt)
        assert np.array_equal(np.asarray(b.hough_points), np.asarray(a.hough_points))
        b.bin_hough_points(64, 48)
        assert np.array_equal(b.hist, a.hist) and np.array_equal(b.xedges, a.xedges)
        assert np.array_equal(b.hough_lines, a.hough_lines)


def test_facade_fit_with_brute_force_hough():
    """test/test_lt_sprat_manual_atlas.py:124 drives the fit from the brute-force transform."""
    from rascal_b200.calibrator import Calibrator, LineList

    g = load_golden("poly_quadratic_deg2")
    cfg = g["config"]
    c = Calibrator(g["peaks"])
    c.set_calibrator_properties(pixel_list=g["pixel_list"], seed=cfg["seed"], **cfg["calib"])
    c.set_hough_properties(**cfg["hough"])
    c.set_ransac_properties(**cfg["ransac"])
    c.set_atlas(LineList(g["lines"]), **cfg["atlas_kw"])
    c.do_hough_transform(brute_force=True)
    pts = O.hough_points_brute_force(c.pairs[:, 0], c.pairs[:, 1], c.min_slope, c.max_slope, c.min_intercept,
                                     c.max_intercept)
    assert len(c.hough_points) == len(pts) and sha(np.asarray(c.hough_points)) == sha(pts)
    hist, xe, ye = O.hough_histogram(pts, c.xbins, c.ybins)
    assert np.array_equal(c.ht.hist, hist) and np.array_equal(c.ht.xedges, xe)
    res = c.fit(max_tries=cfg["fit"].get("max_tries", 500), **{k: v for k, v in cfg["fit"].items() if k != "max_tries"})
    wl = np.polynomial.polynomial.polyval(g["pixel_list"], res[0])
    ref = np.polynomial.polynomial.polyval(g["pixel_list"], g["fit_coeff"])
    assert np.max(np.abs(wl - ref)) < 1.0  # the same wavelength solution from the brute-force transform
