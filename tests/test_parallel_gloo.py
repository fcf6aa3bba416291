"""N>1 host logic on CPU: world_size-2 gloo group, the same code path the NCCL ranks run
(rascal_b200.parallel): id-range sharding, the single all-gather of winner records and the
merge rule `lower cost, later hypothesis id on ties` (calibrator.py:636)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _record(cost, hyp_id, n_inl, counters):
    from rascal_b200 import _native

    r = _native.RansacResult()
    r.cost, r.hyp_id, r.n_match, r.n_inliers = cost, hyp_id, n_inl + 3, n_inl
    for k in range(5):
        r.coeffs[k] = cost * (k + 1)
    for k in range(8):
        r.counters[k] = counters[k]
    return r


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rascal_b200 import parallel

    assert parallel.rank_world(None) == (rank, world)
    assert parallel.rank_world(False) == (0, 1)
    key = parallel.broadcast_int(1234567890123456789 + rank, None)
    lo, hi = parallel.shard_range(1001, rank, world)
    # problem 0: rank 1 holds the lower cost; problem 1: equal costs -> later id wins; problem 2: nobody
    mine = [
        _record(2.0 - rank, 10 + rank, 20 + rank, [hi - lo, 0, 1, 2, 3, 4, 5, 6]),
        _record(0.5, 100 * (rank + 1), 7, [1] * 8),
        _record(float("inf"), -1, 0, [rank] * 8),
    ]
    merged = parallel.allgather_winners(mine, None)
    out = [(m.cost, m.hyp_id, m.n_inliers, list(m.counters), list(m.coeffs[:5])) for m in merged]
    q.put((rank, key, (lo, hi), out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, key0, rng0, out0), (r1, key1, rng1, out1) = got
    assert key0 == key1 == 1234567890123456789  # rank 0's key everywhere
    assert rng0 == (0, 501) and rng1 == (501, 1001)  # disjoint cover
    assert out0 == out1  # identical result on every rank
    p0, p1, p2 = out0
    assert (p0[0], p0[1], p0[2]) == (1.0, 11, 21) and p0[4] == [1.0, 2.0, 3.0, 4.0, 5.0]
    assert p0[3][0] == 1001 and p0[3][1:] == [0, 2, 4, 6, 8, 10, 12]  # counters summed
    assert (p1[0], p1[1]) == (0.5, 200)  # tie: later hypothesis id
    assert p2[1] == -1 and p2[0] == float("inf")


def test_shard_ranges_partition():
    from rascal_b200 import parallel

    for n in (0, 1, 7, 8, 10 ** 8 + 3):
        for world in (1, 2, 3, 4, 8):
            r = [parallel.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_merge_rule():
    from rascal_b200 import parallel

    a = _record(1.0, 5, 9, [1] * 8)
    b = _record(1.0, 9, 9, [2] * 8)
    c = _record(0.9, 2, 9, [3] * 8)
    none = _record(float("inf"), -1, 0, [4] * 8)
    assert parallel.merge_winners([a, b]).hyp_id == 9
    assert parallel.merge_winners([b, a]).hyp_id == 9
    assert parallel.merge_winners([a, b, c, none]).hyp_id == 2
    assert list(parallel.merge_winners([a, b, c, none]).counters) == [10] * 8
    assert parallel.merge_winners([none]).hyp_id == -1


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import types

    from rascal_b200 import parallel

    n_total = 7  # ragged: rank 0 holds 4 spectra, rank 1 holds 3
    mine = []
    for i in range(rank, n_total, world):
        if i == 4:
            mine.append(types.SimpleNamespace(fit_coeff=None))
        else:
            mine.append(types.SimpleNamespace(fit_coeff=np.arange(5) + 10.0 * i, matched_peaks=[0] * (i + 6), rms=0.1 * i,
                                              cost=1e-6 * (i + 1), peak_utilisation=0.5, atlas_utilisation=0.25))
    rec = parallel.gather_records(parallel.result_records(mine), n_total, None)
    ints = parallel.allgather_ints([10 + rank, 7 * rank], None)
    assert np.array_equal(ints, [[10, 0], [11, 7]])
    q.put((rank, rec))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_of_result_records():
    """The final gather of the many-spectra path: spectrum i lives on rank i mod world; afterwards
    every rank holds all records in batch order."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got[0], got[1]) and got[0].shape == (7, 14)
    for i in range(7):
        if i == 4:
            assert got[0][i, 0] == 0 and np.isinf(got[0][i, 3])
        else:
            assert got[0][i, 0] == 1 and got[0][i, 1] == i + 6 and np.isclose(got[0][i, 2], 0.1 * i)
            assert np.array_equal(got[0][i, 6:11], np.arange(5) + 10.0 * i)

