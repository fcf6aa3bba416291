"""Checker of the RANSAC sampling law (calibrator.py:521-566) for arrays of draws
[n, s, 2] = (peak index, candidate index inside the peak's list), -1 rows = abandoned draws.  Used on
the device's draws (tests/test_gpu_sampler.py) and, on the CPU, on NumPy simulations of the law and of
deliberately biased samplers (tests/test_sampler_checker.py), so the checker itself is tested."""
import numpy as np
from scipy import stats

P_FALSE_ALARM = 1e-6


def chi2_ok(observed, expected, what):
    observed, expected = np.asarray(observed, dtype=float).ravel(), np.asarray(expected, dtype=float).ravel()
    keep = expected > 0
    assert observed[~keep].sum() == 0, what + ": events in cells of probability zero"
    o, e = observed[keep], expected[keep]
    small = e < 20  # pool the thin cells into one
    if small.any():
        o, e = np.append(o[~small], o[small].sum()), np.append(e[~small], e[small].sum())
    if len(e) < 2:
        return 0.0
    stat = float(((o - e) ** 2 / e).sum())
    limit = float(stats.chi2.ppf(1 - P_FALSE_ALARM, len(e) - 1))
    assert stat < limit, "%s: chi2 %.1f over %d cells (limit %.1f)" % (what, stat, len(e), limit)
    return stat


def chi2_two_sample(a, b, what):
    a, b = np.asarray(a, dtype=float).ravel(), np.asarray(b, dtype=float).ravel()
    keep = (a + b) > 0
    a, b = a[keep], b[keep]
    k1, k2 = np.sqrt(b.sum() / a.sum()), np.sqrt(a.sum() / b.sum())
    stat = float((((k1 * a - k2 * b) ** 2) / (a + b)).sum())
    limit = float(stats.chi2.ppf(1 - P_FALSE_ALARM, max(len(a) - 1, 1)))
    assert stat < limit, "%s: two-sample chi2 %.1f over %d cells (limit %.1f)" % (what, stat, len(a), limit)
    return stat


def simulate_positions(prob, s, n, rng):
    """Successive sampling without replacement == ranking by log p + Gumbel noise (Plackett-Luce);
    returns the [n, s] matrix of drawn indices, in draw order."""
    out = np.empty((n, s), dtype=np.int64)
    lp = np.log(prob)
    step = 200_000
    for a in range(0, n, step):
        m = min(step, n - a)
        g = lp[None, :] + rng.gumbel(size=(m, len(prob)))
        out[a:a + m] = np.argsort(-g, axis=1)[:, :s]
    return out


def free_entries(setup, pk_k, used):
    """[rows, Kmax] mask of the entries of peak pk_k's candidate array whose wavelength is not in `used`."""
    K = np.diff(setup.cand_off)
    kmax = int(K.max())
    j = np.arange(kmax)[None, :]
    valid = j < K[pk_k][:, None]
    idx = np.minimum(setup.cand_off[pk_k][:, None] + j, len(setup.cand_wid) - 1)
    ent = setup.cand_wid[idx]
    free = valid & ~(ent[:, :, None] == used[:, None, :]).any(axis=2)
    return free, ent


def simulate_law(setup, n, rng, peak_prob=None):
    """NumPy simulation of the whole law: [n, s, 2] draws, -1 rows where the draw is impossible."""
    s = setup.sample_size
    pk = simulate_positions(setup.prob if peak_prob is None else peak_prob, s, n, rng)
    out = np.full((n, s, 2), -1, dtype=np.int32)
    wid = np.full((n, s), -1, dtype=np.int64)
    alive = np.ones(n, dtype=bool)
    ck = np.zeros((n, s), dtype=np.int64)
    for k in range(s):
        free, ent = free_entries(setup, pk[:, k], wid[:, :k])
        n_free = free.sum(axis=1)
        alive &= n_free > 0
        r = np.minimum((rng.random(n) * np.maximum(n_free, 1)).astype(np.int64), np.maximum(n_free - 1, 0))
        pick = np.argmax(np.cumsum(free, axis=1) > r[:, None], axis=1)
        ck[:, k] = pick
        wid[:, k] = ent[np.arange(n), pick]
    out[alive, :, 0] = pk[alive]
    out[alive, :, 1] = ck[alive]
    return out


def check_sample_law(sample, setup, sim_seed=1):
    """Assert that `sample` follows the law of `setup` (engine.RansacSetup); returns the abort fraction."""
    s, prob = setup.sample_size, setup.prob
    n_u = len(setup.peaks)
    n = len(sample)
    assert sample.shape == (n, s, 2)
    aborted = sample[:, 0, 0] < 0
    assert (sample[aborted] == -1).all()
    ok = sample[~aborted]
    pk, ck = ok[:, :, 0].astype(np.int64), ok[:, :, 1].astype(np.int64)
    # --- hard constraints: valid indices, distinct peaks, distinct wavelengths
    assert pk.min() >= 0 and pk.max() < n_u
    srt = np.sort(pk, axis=1)
    assert (srt[:, 1:] != srt[:, :-1]).all(), "a peak was drawn twice"
    K = np.diff(setup.cand_off)
    assert (ck >= 0).all() and (ck < K[pk]).all()
    wid = setup.cand_wid[setup.cand_off[pk] + ck]
    wsrt = np.sort(wid, axis=1)
    assert (wsrt[:, 1:] != wsrt[:, :-1]).all(), "a wavelength was drawn twice"
    # --- the law, against a simulation of it (which also reproduces the abort rate); the first two peak
    #     positions additionally against their exact expectations
    rng = np.random.default_rng(sim_seed)
    sim = simulate_law(setup, n, rng)
    sim_ab = sim[:, 0, 0] < 0
    p_ab = 0.5 * (aborted.mean() + sim_ab.mean())
    assert abs(aborted.mean() - sim_ab.mean()) <= 6.0 * np.sqrt(max(p_ab * (1 - p_ab), 1e-12) * 2.0 / n) + 1e-12, \
        "abort rate %.5f vs the law's %.5f" % (aborted.mean(), sim_ab.mean())
    spk = sim[~sim_ab][:, :, 0].astype(np.int64)
    if not aborted.any() and not sim_ab.any():
        m = len(pk)
        chi2_ok(np.bincount(pk[:, 0], minlength=n_u), m * prob, "first peak")
        joint = np.zeros((n_u, n_u))
        np.add.at(joint, (pk[:, 0], pk[:, 1]), 1)
        expect = m * prob[:, None] * prob[None, :] / (1.0 - prob[:, None])
        np.fill_diagonal(expect, 0.0)
        chi2_ok(joint, expect, "(first, second) peak")
    for k in range(s):
        chi2_two_sample(np.bincount(pk[:, k], minlength=n_u), np.bincount(spk[:, k], minlength=n_u), "peak position %d" % k)
    chi2_two_sample(np.bincount(pk.ravel(), minlength=n_u), np.bincount(spk.ravel(), minlength=n_u), "peak inclusion")
    # --- candidates: the pick is uniform over the entries of the peak's ARRAY whose wavelength is unused.
    #     Without impossible draws that is checked against the exact expectation; where draws can be
    #     abandoned the COMPLETED draws are a biased subset (a first pick that leaves a later peak without a
    #     free wavelength is under-represented), so the same statistics are compared with the simulation's.
    def candidate_stats(pk_, ck_, wid_):
        cells = {}
        rows = np.arange(len(pk_))
        for k in range(s):
            free, _ = free_entries(setup, pk_[:, k], wid_[:, :k])
            assert free[rows, ck_[:, k]].all(), "a used wavelength (or an entry outside the list) was picked"
            n_free = free.sum(axis=1)
            rank = (np.cumsum(free, axis=1) - 1)[rows, ck_[:, k]]
            for nf in np.unique(n_free):
                sel = n_free == nf
                cells.setdefault(int(nf), np.zeros(int(nf)))
                cells[int(nf)] += np.bincount(rank[sel], minlength=int(nf))
        first = np.zeros((n_u, int(K.max())))  # per (peak, entry): duplicated entries must each get their share
        np.add.at(first, (pk_[:, 0], ck_[:, 0]), 1)
        return cells, first

    cells, first = candidate_stats(pk, ck, wid)
    if not aborted.any() and not sim_ab.any():
        n_first = np.bincount(pk[:, 0], minlength=n_u)
        chi2_ok(first, (n_first / K)[:, None] * (np.arange(int(K.max()))[None, :] < K[:, None]), "first candidate")
        for nf, cnt in cells.items():
            if nf > 1:
                chi2_ok(cnt, np.full(nf, cnt.sum() / nf), "candidate rank among %d free entries" % nf)
    else:
        sck = sim[~sim_ab][:, :, 1].astype(np.int64)
        swid = setup.cand_wid[setup.cand_off[spk] + sck]
        scells, sfirst = candidate_stats(spk, sck, swid)
        chi2_two_sample(first, sfirst, "first candidate")
        for nf, cnt in cells.items():
            if nf > 1 and nf in scells:
                chi2_two_sample(cnt, scells[nf], "candidate rank among %d free entries" % nf)
    return float(aborted.mean())
