"""The PRODUCTION sampler of K3 (Philox path), checked on the device against the reference's
sampling law (calibrator.py:521-566):

  * peaks = np.random.choice(peaks, s, replace=False, p=prob): successive sampling - first pick
    ~ prob, second pick given the first ~ prob renormalised over the rest, ... (exact expectations
    for the first two positions, a NumPy simulation of the same law for all of them);
  * one wavelength per chosen peak, uniform over the peak's candidate ARRAY (duplicated entries
    count twice, App. A.3), redrawn while the wavelength is already in the sample - i.e. uniform over
    the entries whose wavelength is still unused;
  * a draw that cannot be completed is abandoned (the try is consumed).

Every draw is read back through rb2_ransac_desc.d_out_sample and handed to tests/sampler_law.py (whose
own power is checked on the CPU by tests/test_sampler_checker.py).  Chi-square statistics are compared
with the 1 - 1e-6 quantile: a correct sampler fails once in a million runs; at 2e6 draws a 10 % error in ONE
peak's probability, a wrong interval shift or a pick over distinct values instead of array entries is far outside."""
import numpy as np
import pytest

from tests.conftest import load_golden
from tests.sampler_law import check_sample_law

pytestmark = pytest.mark.gpu
N_DRAWS = 2_000_000


def draw(setup, n, seed):
    from rascal_b200 import engine

    rb = engine.RansacBatch([setup])
    out = rb.run(0, n, seed=seed, samples=True)
    c = engine.counters_dict(rb.results()[0])
    sample = out["sample"]
    assert (sample != -2).all(), "a hypothesis was never drawn"
    assert c["aborted"] == int((sample[:, 0, 0] < 0).sum()) and c["drawn"] + c["aborted"] == n
    return sample, c


def test_philox_draws_follow_sampling_law():
    """C3 candidate table (60 peaks x 5 entries, 23 peaks with duplicated entries), sample 5 = deg + 1:
    the exactly determined fast path (interval-shift peak sampler, first-try candidates + deferred redraws)."""
    from tests.test_gpu_ransac_refit import setup_from_golden

    s = setup_from_golden(load_golden("c3_deg4"))
    n_distinct = np.array([len(np.unique(s.cand_wave[a:b])) for a, b in zip(s.cand_off[:-1], s.cand_off[1:])])
    assert (np.diff(s.cand_off) > n_distinct).any(), "this case should have duplicated candidate entries"
    sample, c = draw(s, N_DRAWS, seed=20260101)
    assert c["aborted"] == 0
    check_sample_law(sample, s, sim_seed=1)


@pytest.mark.parametrize("case,seed", [("c1_sprat_xe", 11), ("c2_gmos_cuar", 12), ("poly_quadratic_deg2", 13)])
def test_philox_law_on_recorded_tables(case, seed):
    """Real-data tables (LT/SPRAT deg 4, GMOS deg 5 = sample 6) and the deg-2 known-answer case: uneven
    sampling laws, ragged candidate lists, other sample sizes of the compile-time sampler."""
    from tests.test_gpu_ransac_refit import setup_from_golden

    s = setup_from_golden(load_golden(case))
    sample, _ = draw(s, N_DRAWS // 2, seed=seed)
    check_sample_law(sample, s, sim_seed=seed)


def test_philox_law_overdetermined_sample():
    """sample_size 8 > deg + 1: the run-time sampler of the least-squares path (rejection loops)."""
    from tests.test_gpu_ransac_refit import setup_from_golden

    s = setup_from_golden(load_golden("c3_deg4"), sample_size=8)
    sample, _ = draw(s, N_DRAWS // 4, seed=5)
    check_sample_law(sample, s, sim_seed=5)


def test_law_is_the_same_in_every_launch_split():
    """The draw of hypothesis id is a function of (seed, id) only: chunks, grids and deferred rounds
    must not change it."""
    from rascal_b200 import engine
    from tests.test_gpu_ransac_refit import setup_from_golden

    s = setup_from_golden(load_golden("c3_deg4"))
    n = 300_000
    ref = None
    for bpp in (0, 1, 37):
        rb = engine.RansacBatch([s], blocks_per_problem=bpp)
        whole = rb.run(0, n, seed=4, samples=True)["sample"]
        lo = rb.run(0, n // 3, seed=4, samples=True)["sample"]
        hi = rb.run(n // 3, n - n // 3, seed=4, samples=True)["sample"]
        assert np.array_equal(whole, np.concatenate((lo, hi)))
        ref = whole if ref is None else ref
        assert np.array_equal(whole, ref)


def tiny_setup(cands, deg, **kw):
    from rascal_b200 import engine

    x = np.concatenate([[float(100 * (i + 1) + 7 * i * i)] * len(c) for i, c in enumerate(cands)])  # off the histogram edges
    y = np.concatenate([np.asarray(c, dtype=float) for c in cands])
    e = np.linspace(1.0, 2.0, 5)
    return engine.RansacSetup(x, y, fit_deg=deg, sample_size=deg + 1, ransac_tolerance=5.0, min_intercept=-1e9,
                              max_intercept=1e9, pixel_list=np.arange(2048), hist=np.ones((4, 4)), xedges=e,
                              yedges=e + 3000, **kw)


def test_impossible_draws_abort_and_consume_the_try():
    """Peaks 0-2 only know wavelength 5000: a sample holding two of them cannot be completed
    (calibrator.py:557-566).  The abort rate is the law's, the other draws stay valid and uniform."""
    cands = [[5000.0], [5000.0], [5000.0, 5000.0], [5100.0, 5200.0], [5300.0, 5000.0], [5400.0], [5500.0, 5600.0]]
    s = tiny_setup(cands, deg=2)
    sample, c = draw(s, 400_000, seed=99)
    p_ab = check_sample_law(sample, s, sim_seed=3)
    assert 0.02 < p_ab < 0.9


def test_every_draw_impossible():
    """All peaks share one wavelength: every hypothesis aborts, nothing is scored, the kernel returns."""
    from rascal_b200 import engine

    s = tiny_setup([[5000.0]] * 6, deg=1)
    rb = engine.RansacBatch([s])
    out = rb.run(0, 50_000, seed=1, per_hypothesis=True)
    c = engine.counters_dict(rb.results()[0])
    assert c["aborted"] == 50_000 and c["drawn"] == 0 and rb.results()[0].hyp_id == -1
    assert (out["status"] == 0).all()
