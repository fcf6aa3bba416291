"""K3 (batched RANSAC) and K4 (winner refit) through the C ABI vs the oracle and the
golden recordings of the reference's seeded runs.

Tolerances (north_star): identical status / inlier sets / best-model selection on
replayed sample sets; fitted (refit) coefficients within 1e-9 relative of the exact
optimum the reference's TRF iterates towards; RMS within 1e-6 A.  Per-hypothesis
coefficients come from a different (register-resident) solver than LAPACK gelsd and
are compared through the predictions they make on the detector (<= 1e-7 A)."""
import numpy as np
import pytest

from oracle import rascal_oracle as O
from tests.conftest import load_golden
from tests.helpers import problem_from_golden

pytestmark = pytest.mark.gpu
P = np.polynomial.polynomial


def setup_from_golden(g, **over):
    from rascal_b200 import engine

    cfg = g["config"]
    r, f = cfg["ransac"], cfg["fit"]
    lo, hi, min_i, max_i = g["limits"]
    kw = dict(fit_deg=f.get("fit_deg", 4), sample_size=int(g["sample_size"]),
              ransac_tolerance=r.get("ransac_tolerance", 5), min_intercept=min_i, max_intercept=max_i,
              pixel_list=g["pixel_list"], hist=g["hist"].astype(float), xedges=g["xedges"], yedges=g["yedges"],
              hough_weight=r.get("hough_weight", 1.0), filter_close=r.get("filter_close", False),
              minimum_matches=min(r.get("minimum_matches", 0), len(g["lines"]), len(g["peaks"])),
              minimum_peak_utilisation=r.get("minimum_peak_utilisation", 0.0),
              minimum_fit_error=r.get("minimum_fit_error", 1e-4), n_peaks_total=len(g["peaks"]),
              fit_type=f.get("fit_type", "poly"))
    if cfg.get("known"):  # set_known_pairs: rows appended to every sample (calibrator.py:568-573)
        kw.update(pix_known=np.array(cfg["known"][0], dtype=float), wave_known=np.array(cfg["known"][1], dtype=float))
    kw.update(over)
    return engine.RansacSetup(g["candidate_peak"], g["candidate_arc"], **kw)


def ref_curve(g, x, coeffs):
    """The reference's self.polyval on ITS coefficients (basis coefficients for legendre / chebyshev)."""
    return O.BASES[g["config"]["fit"].get("fit_type", "poly")][1](x, coeffs)


def monomial_coeffs(g, coeffs):
    """Recorded hypothesis coefficients as the monomial coefficients the device ABI carries."""
    ft = g["config"]["fit"].get("fit_type", "poly")
    if ft == "poly":
        return np.asarray(coeffs)
    conv = np.polynomial.legendre.leg2poly if ft == "legendre" else np.polynomial.chebyshev.cheb2poly
    c = np.atleast_2d(coeffs)
    out = np.zeros_like(c)
    for i, row in enumerate(c):
        v = conv(row)
        out[i, :len(v)] = v
    return out.reshape(np.shape(coeffs))


def test_setup_matches_reference_tables(golden):
    s = setup_from_golden(golden)
    assert np.array_equal(s.peaks, golden["tr_ransac_peaks"])
    assert np.array_equal(s.cand_wave, golden["tr_cand_flat"])
    assert np.array_equal(np.diff(s.cand_off), golden["tr_cand_len"])


def test_replay_reference_draws(golden):
    """Every sample the seeded reference drew, replayed on the device."""
    from rascal_b200 import engine

    g = golden
    s = setup_from_golden(g)
    rb = engine.RansacBatch([s])
    n = len(g["tr_status"])
    out = rb.run(0, n, replay=(g["tr_x_idx"], g["tr_y_idx"]), per_hypothesis=True)
    # status: identical decisions (intercept / monotonic / scored)
    assert np.array_equal(out["status"], g["tr_status"].astype(np.uint8))
    sc = g["tr_status"] == O.ST_SCORED
    assert sc.sum() > 0
    # hypothesis polynomials.  Interpolation through near-coincident peaks is arbitrarily
    # ill-conditioned (numpy's gelsd and any other solver then differ), so: (i) every exactly
    # determined fit passes through its own sample points, (ii) the scored (well-behaved)
    # hypotheses trace the reference's curve over the detector.
    if s.sample_size == s.deg + 1 and len(s.known_x) == 0:
        xs = s.peaks[g["tr_x_idx"]]
        ys = np.array([[s.cand_wave[s.cand_off[k] + j] for k, j in zip(rx, ry)]
                       for rx, ry in zip(g["tr_x_idx"], g["tr_y_idx"])])
        fit = np.array([P.polyval(x, c) for x, c in zip(xs, out["coeffs"])])
        assert np.max(np.abs(fit - ys)[sc]) < 1e-7
        well = np.abs(g["tr_coeffs"]).max(axis=1) < 1e6
        assert np.max(np.abs(fit - ys)[well]) < 1e-6
    grid = np.linspace(s.pix_min, s.pix_max, 64)
    dev = P.polyval(grid, out["coeffs"][sc].T)
    ref = ref_curve(g, grid, g["tr_coeffs"][sc].T)
    assert np.max(np.abs(dev - ref)) < 1e-6  # incl. the 20 % extrapolation beyond the outer peaks
    assert np.array_equal(out["counts"][sc, 0], g["tr_n_match"][sc])
    assert np.array_equal(out["counts"][sc, 1], g["tr_n_inl"][sc])
    # (atol: empty corner bins give a weight that is pure spline noise, << the 1e-9 the cost adds;
    #  rtol: a pixel inside the spline window turns 1e-12 of coefficient noise into ~1e-11 of weight)
    dw = np.abs(out["weight"][sc] - g["tr_weight"][sc])
    assert np.all(dw <= 1e-10 * np.abs(g["tr_weight"][sc]) + 1e-11), (dw.max(), out["weight"][sc][:3], g["tr_weight"][sc][:3])
    # cost = sum(err)/(n-d)/weight: relative 1e-8, or absolute = 1e-8 A of error per peak
    # (exact synthetic atlases give costs that are pure rounding noise)
    atol = 1e-8 / g["tr_weight"][sc]
    assert np.all(np.abs(out["cost"][sc] - g["tr_cost"][sc]) <= atol + 1e-8 * np.abs(g["tr_cost"][sc]))
    # counters
    c = engine.counters_dict(rb.results()[0])
    assert c["drawn"] == n and c["aborted"] == 0 and c["unsupported"] == 0
    assert c["intercept"] == int((g["tr_status"] == O.ST_INTERCEPT).sum())
    assert c["monotonic"] == int((g["tr_status"] == O.ST_MONOTONIC).sum())
    assert c["scored"] == int(sc.sum())


# measured |coefficient - recording| / |recording| of K4 mode 0 per case (GPUTEST r02 run, printed by the
# test below) - what SciPy's TRF leaves between two implementations of the same iteration;
# the assertion allows 10x the measured gap
COEFF_GAP = {"c3_deg4": 3.6e-7, "poly_linear_deg1": 2.1e-11, "poly_quadratic_deg2": 2e-5, "fine_deg5": 1.2e-7,
             "c1_sprat_xe": 3.1e-7, "c5_deimos": 2.5e-7, "c2_gmos_cuar": 3.7e-9, "c4_npix1024": 1.9e-7,
             "c4_npix2048": 2.6e-8, "c4_npix4096": 5.5e-6, "synth_nm_deg4": 2e-5, "c5_deimos_full": 2e-5, "legendre_deg2": 2e-5,
             "chebyshev_deg2": 2e-5}


def test_replay_selects_reference_winner(golden, request):
    """Best-model selection (App. A.4) + K4: the device winner is the hypothesis the reference accepted
    last - no tie allowance; refit model and inlier set agree with the reference.

    The cost ranks hypotheses by sum(err) in the reference's own summation order (stage C3).  On the exact
    synthetic atlases of the two known-answer cases every good hypothesis is a perfect fit whose errors
    are the rounding noise of the fit itself (sum(err) ~1e-11 A): there the reference's own polyfit
    coefficients are injected (d_replay_coeffs), so that selection is decided on identical numbers."""
    from rascal_b200 import engine

    g = golden
    s = setup_from_golden(g)
    rb = engine.RansacBatch([s])
    n = len(g["tr_status"])
    ref_id = int(np.flatnonzero(g["tr_accepted"])[-1])
    kw = dict(replay=(g["tr_x_idx"], g["tr_y_idx"]))
    err_sum = g["tr_cost"][ref_id] * (g["tr_weight"][ref_id] + 1e-9) * (g["tr_n_match"][ref_id] - s.deg)
    if err_sum < 1e-8:  # the winner's summed error over ~25 peaks is below 1e-8 A: rounding noise of the fit
        kw["replay_coeffs"] = monomial_coeffs(g, g["tr_coeffs"])
    rb.run(0, n, **kw)
    win = rb.results()[0]
    for _ in range(8):
        # a hypothesis that beats the best but fails a refit-dependent acceptance test (here: exact
        # interpolation of deg+1 inliers -> rms < minimum_fit_error, calibrator.py:673-679) is skipped
        # by the reference too: it refitted it and moved on.  K3 is re-run below that (cost, id) floor.
        if win.hyp_id == ref_id or rb.refit()[0][0].accepted:
            break
        assert g["tr_refit_tried"][win.hyp_id] and not g["tr_accepted"][win.hyp_id]
        rb.run(0, n, floors=[(win.cost, win.hyp_id)], **kw)
        win = rb.results()[0]
    assert win.hyp_id == ref_id
    assert win.n_inliers == g["tr_n_inl"][ref_id] and win.n_match == g["tr_n_match"][ref_id]
    out, mx, my, rs = rb.refit()  # mode 0: SciPy's TRF iteration restated on the device
    o = out[0]
    assert o.accepted == 1
    m = o.n_inliers
    assert np.array_equal(mx[0, :m], g["matched_peaks"])
    assert np.array_equal(my[0, :m], g["matched_atlas"])
    deg = s.deg
    coeffs = np.array(o.coeffs[: deg + 1])
    # same iteration path as SciPy on the same inlier set: number of function evaluations
    info = {}
    ref = O.robust_polyfit(g["matched_peaks"], g["matched_atlas"], deg, info=info)
    assert abs(o.huber_iters - info["nfev"]) <= 2
    # TRF stops 1e-8..1e-5 (relative) short of the optimum, chaotically in the last bits of its
    # iterates: device vs this SciPy build agree at that level, and both sit that
    # close to the exact optimum; the wavelength solution itself agrees far below 1e-6 A.
    exact = O.exact_refit(g["matched_peaks"], g["matched_atlas"], deg)
    # coefficients whose true value is zero (linear data fitted at degree 4) have no relative scale: they are
    # held to what they contribute over the detector (1e-5 A; the curves themselves are compared below) and left out
    # of the relative gaps
    atol = 1e-5 / np.max(np.abs(g["matched_peaks"])) ** np.arange(deg + 1)
    sig = np.abs(exact) > 100 * atol

    def rel(a, b):
        return float(np.max(np.abs(a - b)[sig] / np.abs(b)[sig]))

    case = request.node.callspec.id
    print("COEFF_GAP %s: device-recording %.2e  device-exact %.2e  recording-exact %.2e  scipy(now)-recording %.2e"
          % (case, rel(coeffs, g["fit_coeff"]), rel(coeffs, exact), rel(g["fit_coeff"], exact), rel(ref, g["fit_coeff"])))
    gap = COEFF_GAP.get(case, 2e-5)  # cases without a measured entry: the widest measured gap
    assert np.allclose(coeffs, g["fit_coeff"], rtol=gap, atol=atol)
    assert np.allclose(coeffs, exact, rtol=max(gap, 10 * rel(g["fit_coeff"], exact)), atol=atol)
    # (the returned fit_coeff is models.robust_polyfit's MONOMIAL vector for every fit_type)
    assert np.max(np.abs(P.polyval(g["matched_peaks"], coeffs)
                         - P.polyval(g["matched_peaks"], g["fit_coeff"]))) < 1e-6
    assert abs(o.rms - float(g["rms"])) < 1e-6
    assert np.allclose(rs[0, :m], g["residual"], rtol=0, atol=1e-6)
    # mode 1: the exact Huber optimum, 1e-9 relative against the closed-form least squares
    out1, mx1, my1, rs1 = rb.refit(mode=1)
    c1 = np.array(out1[0].coeffs[: deg + 1])
    assert np.allclose(c1, exact, rtol=1e-9, atol=atol * 1e-2)
    assert out1[0].huber_iters == 0  # inliers never leave Huber's quadratic zone
    assert np.array_equal(my1[0, :m], g["matched_atlas"])


def test_replay_with_reference_coefficients_is_decision_exact(golden):
    """With the reference's own polyfit output injected, every later decision runs on identical numbers:
    status, match / inlier counts of every draw, and the cost to the last digits the weight allows."""
    from rascal_b200 import engine

    g = golden
    s = setup_from_golden(g)
    rb = engine.RansacBatch([s])
    n = len(g["tr_status"])
    inj = monomial_coeffs(g, g["tr_coeffs"])
    out = rb.run(0, n, replay=(g["tr_x_idx"], g["tr_y_idx"]), replay_coeffs=inj, per_hypothesis=True)
    assert np.array_equal(out["status"], g["tr_status"].astype(np.uint8))
    sc = g["tr_status"] == O.ST_SCORED
    assert np.array_equal(out["coeffs"][sc], inj[sc])
    assert np.array_equal(out["counts"][sc, 0], g["tr_n_match"][sc])
    assert np.array_equal(out["counts"][sc, 1], g["tr_n_inl"][sc])
    dw = np.abs(out["weight"][sc] - g["tr_weight"][sc])
    assert np.all(dw <= 1e-10 * np.abs(g["tr_weight"][sc]) + 1e-11)
    # cost = sum(err) / (m - deg) / (weight + 1e-9): errors identical, so the cost inherits the weight's 1e-10
    assert np.allclose(out["cost"][sc], g["tr_cost"][sc], rtol=2e-10, atol=0)
    win = rb.results()[0]
    best = np.flatnonzero(sc & (g["tr_n_inl"] >= s.min_inliers) & (g["tr_n_inl"] >= s.min_inliers_frac))
    j = best[np.lexsort((-best, g["tr_cost"][best]))[0]]
    assert win.hyp_id == j  # min cost, latest id on ties - over the reference's own costs


@pytest.mark.parametrize("case", ["c3_deg4", "fine_deg5", "poly_quadratic_deg2"])
def test_replay_block_layout_independent(case):
    from rascal_b200 import engine

    g = load_golden(case)
    s = setup_from_golden(g)
    n = len(g["tr_status"])
    res = []
    for bpp in (1, 3, 64):
        rb = engine.RansacBatch([s], blocks_per_problem=bpp)
        rb.run(0, n, replay=(g["tr_x_idx"], g["tr_y_idx"]))
        r = rb.results()[0]
        res.append((r.cost, r.hyp_id, r.n_inliers, tuple(r.counters)))
    assert res[0] == res[1] == res[2]


def test_philox_hypotheses_agree_with_oracle():
    """Philox-drawn hypotheses (the sampler's law itself: tests/test_gpu_sampler.py): counters add up and
    every decision / count / weight / cost the device reports is what the oracle computes from the
    device's own coefficients."""
    from rascal_b200 import engine

    g = load_golden("c3_deg4")
    s = setup_from_golden(g)
    rb = engine.RansacBatch([s], blocks_per_problem=8)
    n = 200_000
    out = rb.run(0, n, seed=12345, per_hypothesis=True)
    c = engine.counters_dict(rb.results()[0])
    assert c["drawn"] + c["aborted"] == n
    assert c["drawn"] == c["intercept"] + c["monotonic"] + c["empty"] + c["scored"] + c["unsupported"]
    st = out["status"]
    assert (st != 255).all()
    # oracle agrees with the device on every scored hypothesis' cost, given the device's coefficients
    prob = problem_from_golden(g)
    sc = np.flatnonzero(st == O.ST_SCORED)[:200]
    assert len(sc) > 20
    for i in sc:
        c_dev = out["coeffs"][i]
        assert prob.min_intercept <= c_dev[0] <= prob.max_intercept
        assert O.is_monotonic(prob, c_dev)
        err, mx, my = O.match_bijective(prob, c_dev)
        err[err > prob.tol] = prob.tol
        w = O.hough_weight_sum(prob, c_dev)
        cost = sum(err) / (len(err) - len(c_dev) + 1) / (w + 1e-9)
        assert out["counts"][i, 0] == len(err)
        assert out["counts"][i, 1] == int(sum(err < prob.tol))
        assert np.isclose(out["weight"][i], w, rtol=1e-12, atol=0)
        assert np.isclose(out["cost"][i], cost, rtol=1e-9, atol=0)
    # rejected ones are rejected by the oracle too (sample)
    for i in np.flatnonzero(st == O.ST_MONOTONIC)[:100]:
        assert not O.is_monotonic(prob, out["coeffs"][i])
    for i in np.flatnonzero(st == O.ST_INTERCEPT)[:100]:
        c0 = out["coeffs"][i][0]
        assert c0 < prob.min_intercept or c0 > prob.max_intercept


def test_large_call_equals_its_small_parts():
    """A call big enough for large (> 1e7-id) chunks - stage A then reserves queue-A slots a page ahead and leaves
    tagged holes for stage B to skip (ransac.cu, QA_PAGE) - selects the same winner and counts the same
    outcomes as the same id range evaluated in small calls, which use the per-iteration atomic and other
    chunk lengths (rb2_ransac_chunk_plan)."""
    import ctypes as C

    from rascal_b200 import _native, engine

    g = load_golden("c3_deg4")
    s = setup_from_golden(g)
    n, parts = 40_000_000, 8
    lib = _native.load()
    pp = C.c_longlong(0)
    assert lib.rb2_ransac_chunk_plan(1, n, C.byref(pp), None, None) == 0 and pp.value >= 48 * 3 * 148 * 8 * 64
    assert lib.rb2_ransac_chunk_plan(1, n // parts, C.byref(pp), None, None) == 0 and pp.value < 10_000_000
    rb = engine.RansacBatch([s])
    rb.run(0, n, seed=11)
    whole = rb.results()[0]
    best, counters = None, [0] * 8
    for k in range(parts):
        rb.run(k * (n // parts), n // parts, seed=11)
        r = rb.results()[0]
        counters = [x + y for x, y in zip(counters, r.counters)]
        if r.hyp_id >= 0 and (best is None or (r.cost, -r.hyp_id) < (best.cost, -best.hyp_id)):
            best = _native.RansacResult.from_buffer_copy(bytes(r))
    assert list(whole.counters) == counters
    c = engine.counters_dict(whole)
    assert c["drawn"] + c["aborted"] == n
    assert c["intercept"] + c["monotonic"] + c["empty"] + c["scored"] + c["unsupported"] == c["drawn"]
    assert (whole.hyp_id, whole.cost, whole.n_match, whole.n_inliers) == (best.hyp_id, best.cost, best.n_match, best.n_inliers)
    assert list(whole.coeffs) == list(best.coeffs)


def test_philox_sharding_independent():
    """Hypothesis-id ranges are independent of the launch split: best over [0,N) ==
    merge of the bests over [0,N/2) and [N/2,N) (what the multi-GPU path relies on)."""
    from rascal_b200 import engine

    g = load_golden("c3_deg4")
    s = setup_from_golden(g)
    n = 1 << 20
    rb = engine.RansacBatch([s])
    rb.run(0, n, seed=7)
    whole = rb.results()[0]
    rb.run(0, n // 2, seed=7)
    a = rb.results()[0]
    a = (a.cost, a.hyp_id, list(a.counters))
    rb.run(n // 2, n // 2, seed=7)
    b = rb.results()[0]
    b = (b.cost, b.hyp_id, list(b.counters))
    best = min([a, b], key=lambda t: (t[0], -t[1]))
    assert (whole.cost, whole.hyp_id) == (best[0], best[1])
    # the device-side merge the NCCL path runs after its all-gather (two "ranks" = two record sets)
    import ctypes as C

    import torch

    from rascal_b200 import _native

    sets = _native.struct_array(_native.RansacResult, 2)
    rb.run(0, n // 2, seed=7)
    sets[0] = rb.results()[0]
    rb.run(n // 2, n // 2, seed=7)
    sets[1] = rb.results()[0]
    d_sets = engine._upload_structs(sets)
    d_out = torch.empty(C.sizeof(_native.RansacResult), dtype=torch.uint8, device="cuda")
    _native.check(_native.load().rb2_merge_winners(1, 2, d_sets.data_ptr(), d_out.data_ptr(),
                                                   engine.current_stream_ptr()), "rb2_merge_winners")
    merged = engine._download_structs(d_out, _native.RansacResult, 1)[0]
    assert bytes(merged) == bytes(whole)
    assert list(whole.counters) == [x + y for x, y in zip(a[2], b[2])]
    assert whole.hyp_id >= 0 and whole.n_inliers > s.deg


def test_batched_problems_equal_single():
    """Several spectra in one call == each alone."""
    from rascal_b200 import engine

    gs = [load_golden("c3_deg4"), load_golden("c3_deg4")]
    s0 = setup_from_golden(gs[0])
    s1 = setup_from_golden(gs[1], ransac_tolerance=3)
    n = 1 << 16
    rb = engine.RansacBatch([s0, s1], blocks_per_problem=16)
    rb.run(0, n, seed=3)
    both = [(r.cost, r.hyp_id, r.n_inliers) for r in rb.results()]
    for k, s in enumerate((s0, s1)):
        one = engine.RansacBatch([s], blocks_per_problem=5)
        one.run(0, n, seed=3)
        r = one.results()[0]
        assert (r.cost, r.hyp_id, r.n_inliers) == both[k]


@pytest.mark.parametrize("seed", range(3))
def test_refit_random_inlier_sets_vs_scipy(seed):
    """K4 on hand-made winners: noisy / exact / outlier-laden inlier sets, degrees 1-5, vs the
    real scipy.optimize.least_squares call the reference makes."""
    from rascal_b200 import _native, engine

    rng = np.random.default_rng(40 + seed)
    for deg in (1, 2, 3, 4, 5):
        m = int(rng.integers(deg + 3, 60))
        x = np.sort(rng.uniform(10, 4000, m))
        c = np.array([rng.uniform(3500, 6000), rng.uniform(0.8, 2.5), rng.uniform(-5e-5, 5e-5),
                      rng.uniform(-5e-9, 5e-9), rng.uniform(-5e-13, 5e-13), rng.uniform(-1e-17, 1e-17)])[: deg + 1]
        y = P.polyval(x, c) + rng.normal(0, [0.0, 0.3, 1.5][seed], m)
        # a problem whose candidate table is exactly these pairs; the "winner" is the true curve
        hist = np.ones((4, 4))
        e = np.linspace(1.0, 2.0, 5)
        s = engine.RansacSetup(x, y, fit_deg=deg, sample_size=deg + 1, ransac_tolerance=8.0, min_intercept=0,
                               max_intercept=1e5, pixel_list=np.arange(4096), hist=hist, xedges=e,
                               yedges=e + 3000, minimum_fit_error=0.0)
        rb = engine.RansacBatch([s])
        rb.run(0, 0)  # uploads the descriptors
        win = _native.struct_array(_native.RansacResult, 1)
        win[0].cost, win[0].hyp_id = 1.0, 0
        for k in range(deg + 1):
            win[0].coeffs[k] = c[k]
        out, mx, my, rs = rb.refit(win)
        o = out[0]
        assert o.n_inliers == m and o.accepted == 1
        info = {}
        ref = O.robust_polyfit(x, y, deg, info=info)
        dev = np.array(o.coeffs[: deg + 1])
        assert abs(o.huber_iters - info["nfev"]) <= 3, (deg, o.huber_iters, info)
        assert np.max(np.abs(P.polyval(x, dev) - P.polyval(x, ref))) < 5e-6  # TRF's own stopping noise
        assert abs(o.rms - np.sqrt(np.mean((P.polyval(x, ref) - y) ** 2))) < 1e-6


@pytest.mark.parametrize("seed", [9171440914760070181, 5])
def test_winner_at_scale_is_the_oracles_minimum(seed):
    """Min-selection amplifies any rare per-hypothesis glitch: over 2e6 Philox hypotheses the
    winner record's cost must be what the oracle computes from its coefficients, and no scored
    hypothesis may beat it (seed 9171440914760070181 / id 956017 once hit a 32-wide search bug)."""
    from rascal_b200 import engine

    g = load_golden("c3_deg4")
    s = setup_from_golden(g)
    prob = problem_from_golden(g)
    n = 2_000_000
    rb = engine.RansacBatch([s])
    out = rb.run(0, n, seed=seed, per_hypothesis=True)
    win = rb.results()[0]
    sc = np.flatnonzero(out["status"] == O.ST_SCORED)
    el = sc[out["counts"][sc, 1] >= s.min_inliers]
    j = el[np.lexsort((-el, out["cost"][el]))[0]]
    assert (win.hyp_id, win.cost) == (j, out["cost"][j])
    assert np.all(out["weight"][sc] < 4096 * 700.0) and np.all(out["weight"][sc] > 0)
    # production launch (no per-hypothesis outputs) -> same winner, same counters, bit for bit
    rb.run(0, n, seed=seed)
    fast = rb.results()[0]
    assert bytes(fast) == bytes(win)
    assert s.min_inliers <= win.n_inliers and engine.weight_upper_bound(s) >= out["weight"][sc].max()
    for i in list(el[np.argsort(out["cost"][el])][:40]):
        c = out["coeffs"][i]
        err, mx, my = O.match_bijective(prob, c)
        err[err > prob.tol] = prob.tol
        w = O.hough_weight_sum(prob, c)
        cost = sum(err) / (len(err) - len(c) + 1) / (w + 1e-9)
        assert np.isclose(out["weight"][i], w, rtol=1e-12, atol=0)
        assert np.isclose(out["cost"][i], cost, rtol=1e-9, atol=0)
