"""Multi-GPU plumbing: one process per GPU, hypotheses (or spectra) sharded by id
range, ONE exchange per fit.

RANSAC hypotheses are independent given the (tiny, replicated) candidate table, so
rank r evaluates hypothesis ids ``shard_range(n, r, world)`` of the same Philox
stream - the result does not depend on the number of GPUs.  The only collective
is an all-gather of one fixed-size winner record per rank (cost, hypothesis id,
counts, coefficients, counters: 160 bytes) followed by a local merge with the
reference's tie rule (``cost <= best_cost``: the LATER hypothesis wins ties,
calibrator.py:636).  Latency-bound; NCCL over NVLink on GPUs, gloo in CPU tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N


def shard_range(n, rank, world):
    """Contiguous id range [begin, end) of `rank` when n ids are split over `world` ranks."""
    n, rank, world = int(n), int(rank), int(world)
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def rank_world(group=False):
    """(rank, world) of `group`; group=False means "not distributed"."""
    if group is False:
        return 0, 1
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def broadcast_int(value, group=False):
    """Rank 0's 64-bit integer on every rank."""
    if rank_world(group)[1] == 1:
        return int(value)
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(value)], dtype=torch.int64, device=dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return int(t.item())


def allgather_ints(values, group=False):
    """[world, len(values)] int64 array of every rank's integers (host-side bookkeeping, a few bytes)."""
    rank, world = rank_world(group)
    v = np.asarray(values, dtype=np.int64).reshape(1, -1)
    if world == 1:
        return v
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(v[0].copy()).to(dev)
    out = torch.empty(world * t.numel(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape(world, -1)


def allgather_doubles(values, group=False):
    """[world, len(values)] float64 array of every rank's numbers."""
    rank, world = rank_world(group)
    v = np.asarray(values, dtype=np.float64).reshape(1, -1)
    if world == 1:
        return v
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.from_numpy(v[0].copy()).to(dev)
    out = torch.empty(world * t.numel(), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape(world, -1)


def better(cost_a, id_a, cost_b, id_b):
    """True if (cost_a, id_a) beats (cost_b, id_b): lower cost, later id on ties; id < 0 = none."""
    if id_a < 0:
        return False
    if id_b < 0:
        return True
    return cost_a < cost_b or (cost_a == cost_b and id_a > id_b)


def merge_winners(records):
    """Merge per-rank / per-launch RansacResult structs: best winner, summed counters."""
    out = N.RansacResult()
    out.cost, out.hyp_id = float("inf"), -1
    cnt = [0] * 8
    for r in records:
        for k in range(8):
            cnt[k] += int(r.counters[k])
        if better(r.cost, r.hyp_id, out.cost, out.hyp_id):
            C.memmove(C.byref(out), C.byref(r), C.sizeof(N.RansacResult))
    for k in range(8):
        out.counters[k] = cnt[k]
    return out


def allgather_winners(local, group=False, device=None):
    """All-gather one RansacResult per rank and merge.  `local` is a list of structs
    (one per problem); returns the merged list, identical on every rank."""
    import torch
    import torch.distributed as dist

    if group is False or not (dist.is_available() and dist.is_initialized()):
        return list(local)
    world = dist.get_world_size(group)
    if world == 1:
        return list(local)
    n = len(local)
    sz = C.sizeof(N.RansacResult)
    buf = np.empty(n * sz, dtype=np.uint8)
    for i, r in enumerate(local):
        buf[i * sz:(i + 1) * sz] = np.frombuffer(bytes(r), dtype=np.uint8)
    t = torch.from_numpy(buf)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = t.to(device)
    gathered = torch.empty(world * n * sz, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(gathered, t, group=group)
    raw = gathered.cpu().numpy().reshape(world, n, sz)
    merged = []
    for i in range(n):
        recs = [N.RansacResult.from_buffer_copy(raw[w, i].tobytes()) for w in range(world)]
        merged.append(merge_winners(recs))
    return merged


_GATHER = {}


def exchange_device(d_result, n, group=None):
    """The ONE collective of a sharded fit on device-resident winner records: all-gather the `n`
    per-rank records (NCCL, device to device) and merge them in place (rb2_merge_winners: lower cost,
    later hypothesis id on ties, counters add).  `d_result`: contiguous uint8 CUDA tensor of the records."""
    import torch
    import torch.distributed as dist

    from . import engine

    if group is False or not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    need = world * d_result.numel()
    buf = _GATHER.get("buf")
    if buf is None or buf.numel() < need:
        buf = _GATHER["buf"] = torch.empty(need, dtype=torch.uint8, device=d_result.device)
    gathered = buf[:need]
    dist.all_gather_into_tensor(gathered, d_result, group=group)
    rc = N.load().rb2_merge_winners(int(n), world, gathered.data_ptr(), d_result.data_ptr(), engine.current_stream_ptr())
    N.check(rc, "rb2_merge_winners")


# ---- many spectra: spectrum i -> rank i mod world, one final gather of fixed-size records ----

RECORD_DOUBLES = 14  # ok, n_inliers, rms, cost, peak_utilisation, atlas_utilisation, coeffs[8]


def result_records(results, deg_cap=7):
    """Fixed-size (112-byte) records of per-spectrum results (batch.calibrate_batch output) for the final gather."""
    out = np.zeros((len(results), RECORD_DOUBLES))
    for i, r in enumerate(results):
        if r.fit_coeff is None:
            out[i, 3] = np.inf
            continue
        c = np.asarray(r.fit_coeff, dtype=np.float64)[: deg_cap + 1]
        out[i, 0], out[i, 1], out[i, 2], out[i, 3] = 1.0, len(r.matched_peaks), r.rms, r.cost
        out[i, 4], out[i, 5] = r.peak_utilisation, r.atlas_utilisation
        out[i, 6:6 + len(c)] = c
    return out


def gather_records(local, n_total, group=None):
    """All ranks' records in the order of the original batch.  Rank r holds the records of spectra r, r + world,
    ... (`local`: [n_local, RECORD_DOUBLES]); ONE all_gather_into_tensor of equally padded blocks (NCCL device to
    device on GPUs, gloo in the CPU tests), then the interleave.  Identical on every rank."""
    import torch
    import torch.distributed as dist

    local = np.ascontiguousarray(local, dtype=np.float64).reshape(-1, RECORD_DOUBLES)
    rank, world = rank_world(group)
    if world == 1:
        assert len(local) == n_total
        return local
    per = -(-int(n_total) // world)  # ceil: every rank's block is padded to this many records
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    block = torch.zeros((per, RECORD_DOUBLES), dtype=torch.float64, device=dev)
    block[: len(local)] = torch.from_numpy(local).to(dev)
    gathered = torch.empty((world * per, RECORD_DOUBLES), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(gathered, block, group=group)
    g = gathered.cpu().numpy().reshape(world, per, RECORD_DOUBLES)
    out = np.empty((int(n_total), RECORD_DOUBLES))
    for r in range(world):
        idx = np.arange(r, int(n_total), world)
        out[idx] = g[r, : len(idx)]
    return out

