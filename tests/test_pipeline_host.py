"""Host-side pieces of the device-resident many-spectra pipeline (no GPU needed)."""
import ctypes as C

import numpy as np

from rascal_b200 import _native as N
from rascal_b200 import batch, engine, pipeline


def test_struct_dtypes_match_the_c_abi_layout():
    for s in N.STRUCTS:
        dt = np.dtype(s)
        assert dt.itemsize == C.sizeof(s)
        for name, _ in s._fields_:
            assert dt.fields[name][1] == getattr(s, name).offset


def test_spline_operator_is_the_batch_spline_as_a_linear_map():
    rng = np.random.default_rng(0)
    for xb in (8, 50, 100):
        op, brk = pipeline.spline_operator(xb)
        h = rng.integers(0, 900, size=(3, xb)).astype(np.float64)
        gc0, gcn = np.array([1.0, 0.5, 2.0]), np.array([1.6, 3.5, 2.9])
        breaks, coef = batch._spline_tables(h, gc0, gcn)
        step = (gcn - gc0) / (xb - 1)
        got = (op @ h.T).T.reshape(3, -1, 4) / step[:, None, None] ** np.arange(4)[None, None, :]
        scale = np.abs(coef).max(axis=(0, 1))
        assert np.all(np.abs(got - coef) <= 1e-11 * scale)
        assert np.allclose(gc0[:, None] + brk[None, :-1] * step[:, None], breaks[:, :-1], rtol=1e-15)


def test_sampling_cdf_matches_the_reference_law():
    rng = np.random.default_rng(1)
    for _ in range(50):
        peaks = np.sort(rng.uniform(0, 4000, int(rng.integers(6, 90))))
        h = np.histogram(peaks, bins=10)
        prob = 1.0 / h[0][np.digitize(peaks, h[1], right=True) - 1]  # calibrator.py:521-523
        p2, c32 = engine.sampling_cdf32(peaks)
        assert np.allclose(prob / prob.sum(), p2, rtol=1e-14, atol=0)
        assert c32[-1] == 0xFFFFFFFF and np.all(np.diff(c32.astype(np.int64)) > 0)


def test_eligibility_rules():
    from bench import c4_input

    specs = [batch.normalise_spec(c4_input(i)) for i in range(4)]
    assert pipeline.eligible(specs)
    bad = [dict(s) for s in specs]
    bad[2] = dict(bad[2], peaks=bad[2]["peaks"][::-1].copy())  # not increasing
    assert not pipeline.eligible(bad)
    bad = [dict(s) for s in specs]
    bad[1] = dict(bad[1], ransac=dict(bad[1]["ransac"], filter_close=True))
    assert not pipeline.eligible(bad)
    bad = [dict(s) for s in specs]
    bad[3] = dict(bad[3], hough=dict(bad[3]["hough"], xbins=50))
    assert not pipeline.eligible(bad)
    assert not pipeline.eligible([])


def test_chained_results_index_like_one_sequence():
    from rascal_b200 import pipeline

    class Part(list):
        def k3_tables(self, i):
            return ("tables", self[i])

    parts = [Part(["a0", "a1", "a2"]), Part([]), Part(["c0"]), Part(["d0", "d1"])]
    ch = pipeline.ChainedResults(parts)
    flat = ["a0", "a1", "a2", "c0", "d0", "d1"]
    assert len(ch) == 6 and list(ch) == flat
    assert [ch[i] for i in range(6)] == flat and ch[-1] == "d1" and ch[-6] == "a0"
    assert ch.k3_tables(3) == ("tables", "c0") and ch.k3_tables(5) == ("tables", "d1")


def test_batch_parts_policy(monkeypatch):
    from rascal_b200 import batch

    monkeypatch.delenv("RB2_BATCH_PARTS", raising=False)
    assert batch._batch_parts(100) == 1 and batch._batch_parts(511) == 1 and batch._batch_parts(512) == 2
    monkeypatch.setenv("RB2_BATCH_PARTS", "5")
    assert batch._batch_parts(4096) == 5 and batch._batch_parts(10) == 1
    monkeypatch.setenv("RB2_BATCH_PARTS", "junk")
    assert batch._batch_parts(4096) == 2
