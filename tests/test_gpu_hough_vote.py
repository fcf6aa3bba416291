"""K1 (Hough) and K2 (candidate vote) through the C ABI vs the oracle and the
golden recordings of the reference.  Integer/index work: bit-exact."""
import hashlib

import numpy as np
import pytest

from oracle import rascal_oracle as O
from tests.conftest import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu


def hough_problem(g):
    h = g["config"]["hough"]
    lo, hi, min_i, max_i = g["limits"]
    return dict(pair_x=g["pairs"][:, 0], pair_y=g["pairs"][:, 1], min_slope=lo, max_slope=hi,
                min_intercept=min_i, max_intercept=max_i, n_slopes=int(h["num_slopes"]),
                xbins=int(h["xbins"]), ybins=int(h["ybins"]))


def check_against_golden(hb, i, g):
    st = hb.stats[i]
    assert st.n_points == int(g["n_points"])
    assert np.array_equal(hb.hist(i), g["hist"].astype(np.uint32))
    assert np.array_equal(hb.xedges(i), g["xedges"])
    assert np.array_equal(hb.yedges(i), g["yedges"])
    assert np.array_equal(hb.rank(i), O.rank_bins(g["hist"].astype(float)))
    assert st.max_count == g["hist"].max()


def test_hough_golden(golden):
    from rascal_b200 import engine

    hb = engine.HoughBatch([hough_problem(golden)]).run()
    check_against_golden(hb, 0, golden)


def test_hough_points_golden(golden):
    """Lazy hough_points: the reference's exact rows in its exact order."""
    from rascal_b200 import engine

    hb = engine.HoughBatch([hough_problem(golden)]).run()
    pts = hb.points(0)
    assert pts.shape == (int(golden["n_points"]), 2)
    assert hashlib.sha256(np.ascontiguousarray(pts).tobytes()).digest() == golden["points_sha256"].tobytes()
    assert np.array_equal(pts[:64], golden["points_head"])
    assert np.array_equal(pts[-64:], golden["points_tail"])


def test_hough_batched_ragged():
    """All golden cases (different P, S, bin grids) in ONE call."""
    from rascal_b200 import engine

    gs = [load_golden(n) for n in GOLDEN_CASES]
    hb = engine.HoughBatch([hough_problem(g) for g in gs]).run()
    for i, g in enumerate(gs):
        check_against_golden(hb, i, g)


@pytest.mark.parametrize("spc", [1, 7, 64, 5000])
def test_hough_chunking_independent(spc):
    from rascal_b200 import engine

    g = load_golden("fine_deg5")
    hb = engine.HoughBatch([hough_problem(g)]).run(slopes_per_cta=spc)
    check_against_golden(hb, 0, g)


def oracle_hough(p):
    pts = O.hough_points(p["pair_x"], p["pair_y"], p["min_slope"], p["max_slope"], p["n_slopes"],
                         p["min_intercept"], p["max_intercept"])
    hist, xe, ye = O.hough_histogram(pts, p["xbins"], p["ybins"])
    return pts, hist, xe, ye


@pytest.mark.parametrize("seed", range(6))
def test_hough_random_vs_oracle(seed):
    """Random problems incl. negative pixels, tiny grids, S=1, 1 bin, duplicated pairs."""
    from rascal_b200 import engine

    rng = np.random.default_rng(seed)
    probs = []
    for k in range(5):
        P = int(rng.integers(1, 400))
        x = rng.uniform(-200 if k == 1 else 0, 3000, P)
        y = rng.uniform(3000, 9000, P)
        if k == 2:
            x[: P // 2] = x[0]
            y[: P // 2] = y[0]
        S = 1 if k == 3 else int(rng.integers(2, 700))
        probs.append(dict(pair_x=x, pair_y=y, min_slope=float(rng.uniform(0.5, 1.5)),
                          max_slope=float(rng.uniform(1.6, 3.0)), min_intercept=3000.0, max_intercept=4200.0,
                          n_slopes=S, xbins=int(rng.integers(1, 60)), ybins=int(rng.integers(1, 60))))
    hb = engine.HoughBatch(probs).run()
    for i, p in enumerate(probs):
        pts, hist, xe, ye = oracle_hough(p)
        assert hb.stats[i].n_points == len(pts)
        if len(pts) == 0:
            continue
        assert np.array_equal(hb.hist(i), hist.astype(np.uint32)), (seed, i)
        assert np.array_equal(hb.xedges(i), xe) and np.array_equal(hb.yedges(i), ye)
        assert np.array_equal(hb.rank(i), O.rank_bins(hist))
        assert np.array_equal(hb.points(i), pts)


def test_hough_empty_and_no_survivors():
    from rascal_b200 import engine

    probs = [
        dict(pair_x=np.zeros(0), pair_y=np.zeros(0), min_slope=1.0, max_slope=2.0, min_intercept=0.0,
             max_intercept=1.0, n_slopes=10, xbins=4, ybins=5),
        dict(pair_x=np.array([10.0, 20.0]), pair_y=np.array([9000.0, 9500.0]), min_slope=1.0, max_slope=2.0,
             min_intercept=0.0, max_intercept=1.0, n_slopes=10, xbins=4, ybins=5),
        dict(pair_x=np.array([10.0, np.nan]), pair_y=np.array([np.nan, 5.0]), min_slope=1.0, max_slope=2.0,
             min_intercept=-1e9, max_intercept=1e9, n_slopes=10, xbins=4, ybins=5),
    ]
    hb = engine.HoughBatch(probs).run()
    for i in range(3):
        assert hb.stats[i].n_points == 0
        assert hb.hist(i).sum() == 0
        # np.histogram2d of an empty sample: edges over (0, 1)
        assert np.array_equal(hb.xedges(i), np.linspace(0, 1, 5))
        assert np.array_equal(hb.yedges(i), np.linspace(0, 1, 6))
        assert np.array_equal(hb.rank(i), np.arange(20)[::-1])


def test_hough_fine_bins_full_size():
    """Config-5 shape (51x100 pairs, S=2000, 1000x1000 bins) against the oracle, plus the
    size-independent invariants: total count == surviving points == sum of interval lengths."""
    from rascal_b200 import engine

    rng = np.random.default_rng(5)
    peaks = np.sort(rng.uniform(20, 4000, 51))
    lines = np.sort(rng.uniform(6400, 10600, 100))
    pairs = O.make_pairs(peaks, lines)
    lo, hi, mi, ma = O.hough_limits(6500.0, 10500.0, 500, 100, np.arange(4096))
    p = dict(pair_x=pairs[:, 0], pair_y=pairs[:, 1], min_slope=lo, max_slope=hi, min_intercept=mi,
             max_intercept=ma, n_slopes=2000, xbins=1000, ybins=1000)
    hb = engine.HoughBatch([p]).run()
    pts, hist, xe, ye = oracle_hough(p)
    lo_i, hi_i = hb.intervals(0)
    n = int(np.maximum(hi_i - lo_i + 1, 0).sum())
    assert hb.stats[0].n_points == len(pts) == n == int(hb.hist(0).sum())
    assert np.array_equal(hb.hist(0), hist.astype(np.uint32))
    assert np.array_equal(hb.xedges(0), xe) and np.array_equal(hb.yedges(0), ye)
    rank = hb.rank(0)
    assert np.array_equal(rank, O.rank_bins(hist))
    counts = hb.hist(0).ravel()[rank]
    assert np.all(np.diff(counts.astype(np.int64)) <= 0)  # sortedness
    assert np.array_equal(np.sort(rank), np.arange(10 ** 6))  # a permutation


# ---------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------


def vote_settings(g):
    cfg = g["config"]
    h, r, f = cfg["hough"], cfg["ransac"], cfg["fit"]
    sigma = O.candidate_sigma(h.get("range_tolerance", 500), h.get("linearity_tolerance", 100))
    return dict(pairs=g["pairs"], top_n=r.get("top_n_candidate", 5), weighted=r.get("candidate_weighted", True),
                tol=f.get("candidate_tolerance", 2.0), gauss_denom=2 * sigma ** 2 + 1e-9)


def test_vote_golden(golden):
    from rascal_b200 import engine

    hb = engine.HoughBatch([hough_problem(golden)]).run()
    vb = engine.VoteBatch(hb, [vote_settings(golden)]).run()
    cp, ca = vb.candidates(0)
    assert np.array_equal(cp, golden["candidate_peak"])
    assert np.array_equal(ca, golden["candidate_arc"])


def test_vote_batched_ragged():
    from rascal_b200 import engine

    gs = [load_golden(n) for n in GOLDEN_CASES]
    hb = engine.HoughBatch([hough_problem(g) for g in gs]).run()
    vb = engine.VoteBatch(hb, [vote_settings(g) for g in gs]).run()
    for i, g in enumerate(gs):
        cp, ca = vb.candidates(i)
        assert np.array_equal(cp, g["candidate_peak"]) and np.array_equal(ca, g["candidate_arc"])


@pytest.mark.parametrize("weighted", [True, False])
@pytest.mark.parametrize("seed", range(4))
def test_vote_random_vs_oracle(seed, weighted):
    """Small random spectra: device vote == oracle's literal B*P loop, incl. aggregated
    weights (to rounding) and hit counts (exact)."""
    _vote_random_case(seed, weighted, triple=False)


def test_vote_heavily_duplicated_atlas():
    """Every atlas line three times: more pairs per peak than the flattened work-unit tables hold
    (distinct wavelengths + 64), so the kernel takes its pair-per-thread path; same answers."""
    _vote_random_case(7, True, triple=True)


def test_vote_unsorted_atlas_takes_the_enumeration_path():
    """Atlas lines in random order: pair wavelengths are not ascending inside a peak, so the ordered hit
    list cannot be read off the rank order by bisection and the kernel falls back to the second window
    enumeration; the answers (which depend on the pair order, App. A.3) still equal the oracle's."""
    for seed in (0, 2):
        _vote_random_case(seed, True, triple=False, shuffle=True)


def _vote_random_case(seed, weighted, triple, shuffle=False):
    from rascal_b200 import engine

    rng = np.random.default_rng(100 + seed)
    n_p, n_a = int(rng.integers(5, 25)), int(rng.integers(8, 60))
    if triple:
        n_a = 45
    peaks = np.sort(rng.uniform(10, 1000, n_p))
    truth = 4000 + 2.0 * peaks + 1e-4 * peaks ** 2
    lines = np.sort(np.concatenate((truth[: n_p * 2 // 3], rng.uniform(3800, 6400, n_a))))
    if seed == 1:
        lines = np.concatenate((lines, lines[:3]))  # duplicated atlas wavelengths
    if triple:
        lines = np.concatenate((lines, lines, lines))
        assert len(lines) > len(np.unique(lines)) + 64
    if shuffle:
        lines = rng.permutation(lines)
    pairs = O.make_pairs(peaks, lines)
    lo, hi, mi, ma = O.hough_limits(4000.0, 6100.0, 200, 50, np.arange(1024))
    xb, yb = int(rng.integers(10, 50)), int(rng.integers(10, 50))
    p = dict(pair_x=pairs[:, 0], pair_y=pairs[:, 1], min_slope=lo, max_slope=hi, min_intercept=mi,
             max_intercept=ma, n_slopes=300, xbins=xb, ybins=yb)
    tol = float(rng.uniform(1.0, 12.0))
    sigma = O.candidate_sigma(200, 50)
    top_n = int(rng.integers(1, 7))
    hb = engine.HoughBatch([p]).run()
    vb = engine.VoteBatch(hb, [dict(pairs=pairs, top_n=top_n, weighted=weighted, tol=tol,
                                    gauss_denom=2 * sigma ** 2 + 1e-9)], diagnostics=True).run()
    cp, ca = vb.candidates(0)
    hist = hb.hist(0).astype(float)
    hl = O.hough_lines(hist, hb.xedges(0), hb.yedges(0), O.rank_bins(hist))
    cands = O.candidate_points_linear(hl, pairs, tol, sigma)
    op, oa = O.most_common_candidates(cands, top_n, weighted)
    assert np.array_equal(cp, np.array(op)) and np.array_equal(ca, np.array(oa))
    # diagnostics: hit counts per (peak, wavelength) exact, weights to rounding
    agg, cnt = vb.aggregates(0)
    upeaks, _, _, _, _, uwaves = vb.layouts[0]
    pk = np.concatenate([c[0] for c in cands])
    wv = np.concatenate([c[1] for c in cands])
    pr = np.concatenate([c[2] for c in cands])
    ocnt = np.zeros_like(cnt)
    oagg = np.zeros_like(agg)
    np.add.at(ocnt, (np.searchsorted(upeaks, pk), np.searchsorted(uwaves, wv)), 1)
    np.add.at(oagg, (np.searchsorted(upeaks, pk), np.searchsorted(uwaves, wv)), pr if weighted else 1.0)
    assert np.array_equal(cnt, ocnt)
    assert np.allclose(agg, oagg, rtol=1e-12, atol=0)
