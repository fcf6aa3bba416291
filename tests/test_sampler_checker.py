"""The sampling-law checker (tests/sampler_law.py) is itself test infrastructure for the device
sampler: it must accept draws that follow the law and reject samplers with the defects a device
implementation could have (no GPU needed)."""
import numpy as np
import pytest

from tests.conftest import load_golden
from tests.sampler_law import check_sample_law, simulate_law


def _setup():
    from tests.test_gpu_ransac_refit import setup_from_golden

    return setup_from_golden(load_golden("c3_deg4"))


N = 300_000


def test_checker_accepts_the_law():
    s = _setup()
    check_sample_law(simulate_law(s, N, np.random.default_rng(7)), s, sim_seed=8)


def test_checker_accepts_the_law_with_abandoned_draws():
    """A table where draws can be impossible: completed draws are a biased subset, the checker compares with
    the simulation of the same law instead of the unconditional expectations."""
    from tests.test_gpu_ransac_refit import setup_from_golden

    s = setup_from_golden(load_golden("poly_quadratic_deg2"))
    d = simulate_law(s, N, np.random.default_rng(17))
    assert (d[:, 0, 0] < 0).any()
    check_sample_law(d, s, sim_seed=18)


def test_checker_rejects_a_biased_peak_law():
    s = _setup()
    p = s.prob.copy()
    p[3] *= 1.2  # one peak's interval 20 % too wide (detectable at this N; 5 % needs the GPU test's 2e6 draws)
    p /= p.sum()
    with pytest.raises(AssertionError):
        check_sample_law(simulate_law(s, N, np.random.default_rng(7), peak_prob=p), s, sim_seed=8)


def test_checker_rejects_sampling_with_the_wrong_renormalisation():
    """Second pick drawn from the ORIGINAL law and merely re-drawn on a repeat is the right law; drawing
    it from a law that forgets to remove the first pick's mass (uniform shift) is not."""
    s = _setup()
    rng = np.random.default_rng(9)
    d = simulate_law(s, N, rng)
    # corrupt: swap positions 0 and 1 of every draw whose first peak is below the median (order bias)
    sel = d[:, 0, 0] < len(s.peaks) // 2
    d[sel, 0], d[sel, 1] = d[sel, 1].copy(), d[sel, 0].copy()
    with pytest.raises(AssertionError):
        check_sample_law(d, s, sim_seed=8)


def test_checker_rejects_candidates_picked_from_the_distinct_values():
    """Uniform over the DISTINCT wavelengths of a peak instead of its candidate array (App. A.3:
    duplicated entries bias the draw) must be caught."""
    s = _setup()
    rng = np.random.default_rng(11)
    d = simulate_law(s, N, rng)
    off, wid = s.cand_off, s.cand_wid
    pk = d[:, 0, 0]
    for u in range(len(s.peaks)):
        ent = wid[off[u]:off[u + 1]]
        firsts = np.array([int(np.flatnonzero(ent == w)[0]) for w in np.unique(ent)])
        if len(firsts) < len(ent):
            rows = np.flatnonzero(pk == u)
            d[rows, 0, 1] = firsts[rng.integers(len(firsts), size=len(rows))]
    # repair later picks that now collide, so that only the frequency defect is left
    w_all = wid[off[d[:, :, 0]] + d[:, :, 1]]
    clash = (np.sort(w_all, axis=1)[:, 1:] == np.sort(w_all, axis=1)[:, :-1]).any(axis=1)
    d = d[~clash]
    with pytest.raises(AssertionError):
        check_sample_law(d, s, sim_seed=8)


def test_checker_rejects_a_repeated_wavelength():
    s = _setup()
    d = simulate_law(s, 20_000, np.random.default_rng(5))
    # force a wavelength clash in one draw: give pick 1 the wavelength id of pick 0 if that peak lists it
    wid, off = s.cand_wid, s.cand_off
    done = False
    for r in range(len(d)):
        w0 = wid[off[d[r, 0, 0]] + d[r, 0, 1]]
        ent = wid[off[d[r, 1, 0]]:off[d[r, 1, 0] + 1]]
        hit = np.flatnonzero(ent == w0)
        if len(hit):
            d[r, 1, 1] = hit[0]
            done = True
            break
    assert done
    with pytest.raises(AssertionError, match="wavelength was drawn twice"):
        check_sample_law(d, s, sim_seed=8)
