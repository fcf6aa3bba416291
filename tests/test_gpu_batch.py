"""Many spectra per call (rascal_b200.batch): every spectrum of a ragged batch gets what the
single-spectrum pipeline gives it (same Philox key => same winner)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_batch_equals_single_spectrum_pipeline():
    from bench import c4_input
    from rascal_b200 import batch, engine

    specs = [c4_input(i) for i in range(12)]
    n_hyp, seed = 20_000, 77
    res = batch.calibrate_batch(specs, n_hypotheses=n_hyp, seed=seed)
    assert len(res) == len(specs)
    n_fit = 0
    for i, sp in enumerate(specs):
        ns = batch.normalise_spec(sp)
        lo, hi, mi, ma = ns["limits"]
        h = ns["hough"]
        pairs = np.column_stack((np.repeat(ns["peaks"], len(ns["lines"])), np.tile(ns["lines"], len(ns["peaks"]))))
        hb = engine.HoughBatch([dict(pair_x=pairs[:, 0], pair_y=pairs[:, 1], min_slope=lo, max_slope=hi,
                                     min_intercept=mi, max_intercept=ma, n_slopes=h["num_slopes"],
                                     xbins=h["xbins"], ybins=h["ybins"])]).run()
        sigma = (h["range_tolerance"] + h["linearity_tolerance"]) * 1.1775
        vb = engine.VoteBatch(hb, [dict(pairs=pairs, top_n=5, weighted=True, tol=10.0,
                                        gauss_denom=2 * sigma ** 2 + 1e-9)]).run()
        cp, ca = vb.candidates(0)
        s = engine.RansacSetup(cp, ca, fit_deg=4, sample_size=5, ransac_tolerance=5, min_intercept=mi,
                               max_intercept=ma, pixel_list=np.arange(ns["n_pix"], dtype=float), hist=hb.hist(0).astype(float),
                               xedges=hb.xedges(0), yedges=hb.yedges(0), n_peaks_total=len(ns["peaks"]))
        rb = engine.RansacBatch([s])
        rb.run(0, n_hyp, seed=seed)
        w = rb.results()[0]
        fit, mx, my, rs = rb.refit()
        r = res[i]
        assert r.counters == engine.counters_dict(w)
        if w.hyp_id < 0 or not fit[0].accepted:
            assert r.fit_coeff is None
            continue
        n_fit += 1
        assert np.isclose(r.cost, w.cost, rtol=1e-12)
        m = fit[0].n_inliers
        assert np.array_equal(r.matched_peaks, mx[0, :m]) and np.array_equal(r.matched_atlas, my[0, :m])
        assert np.allclose(r.fit_coeff, np.array(fit[0].coeffs[:5]), rtol=1e-6)
        assert abs(r.rms - fit[0].rms) < 1e-9
        assert r.peak_utilisation == m / len(ns["peaks"])
    assert n_fit >= 8


def test_batch_quality_on_c4_sample():
    """Success criterion of BASELINE configs[3] on a sample: fraction of spectra whose solution is
    within 1 A RMS of the truth over the detector (reported, and sanity-bounded from below)."""
    from bench import c4_input
    from rascal_b200 import batch

    P = np.polynomial.polynomial
    specs = [c4_input(i) for i in range(64)]
    res = batch.calibrate_batch(specs, n_hypotheses=10_000, seed=1)
    good = 0
    for sp, r in zip(specs, res):
        if r.fit_coeff is None:
            continue
        x = np.linspace(0, sp["num_pix"] - 1, 256)
        good += np.sqrt(np.mean((P.polyval(x, r.fit_coeff) - P.polyval(x, sp["truth"])) ** 2)) < 1.0
    print("within 1 A of truth: %d / %d" % (good, len(specs)))
    assert sum(r.fit_coeff is not None for r in res) >= 48


def test_c4_batch_against_reference_recordings():
    """The many-spectra path on ITS OWN inputs against the reference: three spectra of the C4 batch (one per
    detector size; tests/golden/make_golden_c4.py recorded them from the unmodified reference) go through
    calibrate_batch's device pipeline together with neighbours of the batch; the K3 tables the device
    built from its own K1 / K2 outputs (peaks with candidates, candidate lists in order) are the
    reference's, bit for bit, and replaying the reference's recorded draws on those tables gives its
    decisions and its winner."""
    from bench import c4_input
    from rascal_b200 import batch, engine, pipeline
    from tests.conftest import load_golden
    from tests.test_gpu_ransac_refit import setup_from_golden

    gold = [load_golden("c4_npix%d" % n) for n in (1024, 2048, 4096)]
    idx = [int(g["c4_index"]) for g in gold]
    others = [i for i in range(max(idx) + 4) if i not in idx][:9]
    order = idx + others  # ragged batch of 12: the recorded three first
    specs = [batch.normalise_spec(c4_input(i)) for i in order]
    assert pipeline.eligible(specs)
    out = pipeline.run(specs, 5_000, seed=3)
    for k, g in enumerate(gold):
        assert np.array_equal(specs[k]["peaks"], np.sort(g["peaks"])) and np.array_equal(specs[k]["lines"], g["lines"])
        t = out.k3_tables(k)
        assert np.array_equal(t["peaks"], g["tr_ransac_peaks"])
        assert np.array_equal(np.diff(t["cand_off"]), g["tr_cand_len"])
        assert np.array_equal(t["cand_wave"], g["tr_cand_flat"])
        # the device-prepared sampling law / weight tables equal the host set-up's on the recorded inputs
        s = setup_from_golden(g)
        assert np.array_equal(t["cdf32"], s.cdf32)
        d = t["desc"]
        assert (float(d["pix_min"]), float(d["pix_max"])) == (s.pix_min, s.pix_max)
        assert np.allclose(t["pp_coef"], s.pp_coef, rtol=1e-9, atol=1e-9 * np.abs(s.pp_coef).max())
        # replay of the reference's draws on tables identical to the device's
        rb = engine.RansacBatch([s])
        n = len(g["tr_status"])
        st = rb.run(0, n, replay=(g["tr_x_idx"], g["tr_y_idx"]), per_hypothesis=True)
        assert np.array_equal(st["status"], g["tr_status"].astype(np.uint8))
        assert rb.results()[0].hyp_id == int(np.flatnonzero(g["tr_accepted"])[-1])
