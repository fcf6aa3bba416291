"""Device-resident many-spectra pipeline (rascal_b200.pipeline): the tables rb2_ransac_prepare
builds on the device are the ones the host set-up (calibrator.py:441-523 restated in
batch.ransac_setups) builds, and the end-to-end results equal the staged host path's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def c4(n):
    from bench import c4_input

    return [c4_input(i) for i in range(n)]


def test_expand_pairs_is_the_cartesian_product():
    import ctypes as C

    import torch

    from rascal_b200 import _native as N
    from rascal_b200 import engine

    rng = np.random.default_rng(3)
    peaks, lines = np.sort(rng.uniform(0, 4000, 37)), np.sort(rng.uniform(3000, 9000, 211))
    dp, dl = torch.from_numpy(peaks).cuda(), torch.from_numpy(lines).cuda()
    P = len(peaks) * len(lines)
    px, py = torch.empty(P, dtype=torch.float64, device="cuda"), torch.empty(P, dtype=torch.float64, device="cuda")
    wid = torch.empty(P, dtype=torch.int32, device="cuda")
    off = torch.empty(len(peaks) + 1, dtype=torch.int32, device="cuda")
    h = N.struct_array(N.PairsDesc, 1)
    h[0].d_peaks, h[0].d_lines, h[0].n_peaks, h[0].n_lines = dp.data_ptr(), dl.data_ptr(), len(peaks), len(lines)
    h[0].d_pair_x, h[0].d_pair_y, h[0].d_pair_wid, h[0].d_pair_off = px.data_ptr(), py.data_ptr(), wid.data_ptr(), off.data_ptr()
    d = engine._upload_structs(h)
    N.check(N.load().rb2_expand_pairs(1, d.data_ptr(), C.addressof(h), engine.current_stream_ptr()), "rb2_expand_pairs")
    assert np.array_equal(px.cpu().numpy(), np.repeat(peaks, len(lines)))  # itertools.product order (calibrator.py:103)
    assert np.array_equal(py.cpu().numpy(), np.tile(lines, len(peaks)))
    assert np.array_equal(wid.cpu().numpy(), np.tile(np.arange(len(lines)), len(peaks)))
    assert np.array_equal(off.cpu().numpy(), np.arange(len(peaks) + 1) * len(lines))


def test_device_prepared_tables_equal_host_setup():
    from rascal_b200 import batch, engine, pipeline

    raw = c4(20)
    specs = [batch.normalise_spec(sp) for sp in raw]
    assert pipeline.eligible(specs)
    out = pipeline.run(specs, 2000, seed=3)
    hp, vs = [], []
    for sp in specs:
        lo, hi, mi, ma = sp["limits"]
        h = sp["hough"]
        hp.append(dict(pair_x=np.repeat(sp["peaks"], len(sp["lines"])), pair_y=np.tile(sp["lines"], len(sp["peaks"])),
                       min_slope=lo, max_slope=hi, min_intercept=mi, max_intercept=ma, n_slopes=int(h["num_slopes"]),
                       xbins=int(h["xbins"]), ybins=int(h["ybins"])))
        sigma = (h["range_tolerance"] + h["linearity_tolerance"]) * 1.1775
        vs.append(dict(pairs=None, peaks=sp["peaks"], lines=sp["lines"], top_n=sp["ransac"]["top_n_candidate"],
                       weighted=sp["ransac"]["candidate_weighted"], tol=sp["fit"]["candidate_tolerance"],
                       gauss_denom=2 * sigma ** 2 + 1e-9))
    hough = engine.HoughBatch(hp).run()
    vote = engine.VoteBatch(hough, vs).run()
    setups = batch.ransac_setups(specs, hough, vote)
    for i, s in enumerate(setups):
        t = out.k3_tables(i)
        d = t["desc"]
        assert bool(out.enough[i]) == bool(s.enough)
        assert np.array_equal(t["peaks"], s.peaks) and np.array_equal(t["cand_off"], s.cand_off)
        assert np.array_equal(t["cand_wave"], s.cand_wave) and np.array_equal(t["cand_wid"], s.cand_wid)
        assert int(d["n_wids"]) == len(s.uwaves)
        if s.enough:
            assert np.array_equal(t["cdf32"], s.cdf32), i
        assert d["pix_min"] == s.pix_min and d["pix_max"] == s.pix_max
        assert (d["gc_min"], d["gc_max"], d["ic_min"]) == (s.gc_min, s.gc_max, s.ic_min)
        assert (d["h_lo"], d["h_hi"]) == (s.h_lo, s.h_hi)
        assert np.allclose(t["pp_breaks"], s.pp_breaks, rtol=1e-14, atol=0)
        scale = np.abs(np.asarray(s.pp_coef)).reshape(-1, 4).max(axis=0)
        assert np.all(np.abs(t["pp_coef"].reshape(-1, 4) - np.asarray(s.pp_coef).reshape(-1, 4)) <= 1e-10 * scale)
        assert np.isclose(d["w_upper"], engine.weight_upper_bound(s), rtol=1e-9)


def test_device_pipeline_equals_staged_path():
    from rascal_b200 import batch

    raw = c4(24)
    a = batch.calibrate_batch(raw, n_hypotheses=20_000, seed=5, match_tolerance=5.0, device_pipeline=True)
    b = batch.calibrate_batch(raw, n_hypotheses=20_000, seed=5, match_tolerance=5.0, device_pipeline=False)
    assert type(a).__name__ == "BatchResults" and len(a) == len(b) == len(raw)
    n_fit = 0
    for ra, rb in zip(a, b):
        assert ra.counters == rb.counters
        assert (ra.fit_coeff is None) == (rb.fit_coeff is None)
        if ra.fit_coeff is None:
            continue
        n_fit += 1
        assert np.isclose(ra.cost, rb.cost, rtol=1e-12)
        assert np.array_equal(ra.matched_peaks, rb.matched_peaks) and np.array_equal(ra.matched_atlas, rb.matched_atlas)
        assert np.allclose(ra.fit_coeff, rb.fit_coeff, rtol=1e-6) and abs(ra.rms - rb.rms) < 1e-9
        assert ra.peak_utilisation == rb.peak_utilisation and ra.valid == rb.valid
        assert (ra.match is None) == (rb.match is None)
        if ra.match is not None:
            assert np.array_equal(ra.match.matched_atlas, rb.match.matched_atlas)
            assert abs(ra.match.rms - rb.match.rms) < 1e-9
    assert n_fit >= 16


def test_ineligible_batches_take_the_staged_path():
    from rascal_b200 import batch

    raw = c4(4)
    raw[1] = dict(raw[1], ransac=dict(raw[1]["ransac"], filter_close=True))
    res = batch.calibrate_batch(raw, n_hypotheses=5000, seed=1)
    assert isinstance(res, list) and len(res) == 4


def test_large_batch_in_parts_equals_one_part(monkeypatch):
    """calibrate_batch cuts batches of >= 512 spectra into consecutively launched parts (host packing
    overlaps the device); every spectrum gets exactly what the undivided batch gives it."""
    from bench import c4_input
    from rascal_b200 import batch

    specs = [c4_input(i) for i in range(520)]
    monkeypatch.setenv("RB2_BATCH_PARTS", "1")
    one = batch.calibrate_batch(specs, n_hypotheses=2000, seed=3, match_tolerance=5.0)
    monkeypatch.setenv("RB2_BATCH_PARTS", "3")
    tm = {}
    many = batch.calibrate_batch(specs, n_hypotheses=2000, seed=3, match_tolerance=5.0, timings=tm)
    assert tm["parts"] == 3 and len(many) == len(one) == 520
    for i, (a, b) in enumerate(zip(one, many)):
        assert (a.fit_coeff is None) == (b.fit_coeff is None), i
        assert a.cost == b.cost and a.counters == b.counters, i
        if a.fit_coeff is not None:
            assert np.array_equal(a.fit_coeff, b.fit_coeff) and np.array_equal(a.matched_atlas, b.matched_atlas), i
            assert (a.match is None) == (b.match is None)
            if a.match is not None:
                assert np.array_equal(a.match.fit_coeff, b.match.fit_coeff)
    t = many.k3_tables(519)
    assert t["peaks"].shape[0] == int(t["desc"]["n_upeaks"])
    assert many[-1].cost == one[-1].cost
