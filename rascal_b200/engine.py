"""Batched device engine: host-side packing + the C-ABI calls (include/rascal_b200.h).

This is the layer between the reference-shaped facade (``calibrator.py`` /
``houghtransform.py``) and ``librascal_b200.so``.  Every stage takes a *batch* of
problems (leading ``n_problems`` dimension of the C ABI); a single spectrum is a
batch of one.  PyTorch is used for device buffers and streams only.

There is no CPU fallback: every entry point raises ``NativeError`` when the
library is missing or no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

_ALIGN = 32

# bytes moved across PCIe by this module (bench.py reads these for its e2e line)
TRANSFER = {"h2d": 0, "d2h": 0}


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise N.NativeError("rascal_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def current_stream_ptr():
    return int(_torch().cuda.current_stream().cuda_stream)


class Arena:
    """One contiguous device buffer assembled on the host.

    ``add(arr)`` stages input data, ``reserve(nbytes)`` leaves room for an
    output; ``upload()`` moves the staged bytes in ONE host->device copy (from
    pinned memory) and returns the device base address.  Views of regions come
    back as torch tensors / numpy arrays.
    """

    def __init__(self):
        self._chunks = []  # (offset, ndarray)
        self.size = 0
        self.buf = None
        self.host = None
        self.host_cache = None
        self.h2d_bytes = 0

    def _bump(self, nbytes):
        off = self.size
        self.size = (off + int(nbytes) + _ALIGN - 1) // _ALIGN * _ALIGN
        return off

    def add(self, arr):
        a = np.ascontiguousarray(arr)
        off = self._bump(max(a.nbytes, 1))
        self._chunks.append((off, a))
        return off

    def reserve(self, nbytes):
        return self._bump(max(int(nbytes), 1))

    def prefetch(self):
        """ONE device->host copy of the whole arena; later numpy() calls read the host copy."""
        self.host_cache = self.buf.cpu().numpy()
        TRANSFER["d2h"] += self.host_cache.nbytes

    def upload(self, pinned=True):
        torch = _torch()
        n = max(self.size, _ALIGN)
        staged = max((o + a.nbytes for o, a in self._chunks), default=0)
        self.buf = torch.empty(n, dtype=torch.uint8, device="cuda")
        if staged:
            host = torch.empty(staged, dtype=torch.uint8, pin_memory=pinned)
            hv = host.numpy()
            for off, a in self._chunks:
                hv[off:off + a.nbytes] = a.reshape(-1).view(np.uint8)
            self.buf[:staged].copy_(host, non_blocking=True)
            self.host = host  # keep alive until the copy has run
            self.h2d_bytes = staged
            TRANSFER["h2d"] += staged
        return self.ptr(0)

    def ptr(self, off):
        return self.buf.data_ptr() + off

    def tensor(self, off, dtype, count):
        torch = _torch()
        esz = torch.empty(0, dtype=dtype).element_size()
        return self.buf[off:off + count * esz].view(dtype)

    def numpy(self, off, np_dtype, count):
        nbytes = np.dtype(np_dtype).itemsize * count
        if self.host_cache is not None:
            return self.host_cache[off:off + nbytes].view(np_dtype)
        TRANSFER["d2h"] += nbytes
        return self.buf[off:off + nbytes].cpu().numpy().view(np_dtype)


def _upload_structs(arr):
    """ctypes struct array -> device uint8 tensor."""
    torch = _torch()
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    TRANSFER["h2d"] += host.numel()
    return host.cuda()


def _download_structs(tensor, cls, n):
    TRANSFER["d2h"] += tensor.numel()
    raw = tensor.cpu().numpy().tobytes()
    return (cls * n).from_buffer_copy(raw)


# ---------------------------------------------------------------------------
# K1  Hough accumulation
# ---------------------------------------------------------------------------


class HoughBatch:
    """Hough accumulate + bin + rank for a batch of spectra (rb2_hough_accumulate).

    problems: sequence of dicts with keys pair_x, pair_y (float64 arrays),
    min_slope, max_slope, min_intercept, max_intercept, n_slopes, xbins, ybins.
    """

    def __init__(self, problems):
        self.lib = N.load()
        _torch()
        self.n = len(problems)
        self.problems = problems
        ar, wk, sm = Arena(), Arena(), Arena()  # inputs (H2D) | device-only work | small outputs
        self.arena, self.work, self.small = ar, wk, sm
        self.h_desc = N.struct_array(N.HoughDesc, self.n)
        self._off = []
        for i, p in enumerate(problems):
            x = np.asarray(p["pair_x"], dtype=np.float64)
            y = np.asarray(p["pair_y"], dtype=np.float64)
            assert x.shape == y.shape and x.ndim == 1
            P, xb, yb = x.size, int(p["xbins"]), int(p["ybins"])
            B = xb * yb
            o = dict(x=ar.add(x), y=ar.add(y), lo=wk.reserve(4 * P), hi=wk.reserve(4 * P),
                     xe=sm.reserve(8 * (xb + 1)), ye=sm.reserve(8 * (yb + 1)), hist=wk.reserve(4 * B),
                     rank=wk.reserve(4 * B), rankpos=wk.reserve(4 * B), tmp=wk.reserve(4 * B))
            self._off.append(o)
        self.o_stats = sm.reserve(C.sizeof(N.HoughStats) * self.n)
        base, wbase, sbase = ar.upload(), wk.upload(), sm.upload()
        for i, p in enumerate(problems):
            o = self._off[i]
            d = self.h_desc[i]
            d.min_slope, d.max_slope = float(p["min_slope"]), float(p["max_slope"])
            d.min_intercept, d.max_intercept = float(p["min_intercept"]), float(p["max_intercept"])
            d.n_slopes, d.xbins, d.ybins = int(p["n_slopes"]), int(p["xbins"]), int(p["ybins"])
            d.n_pairs = int(np.asarray(p["pair_x"]).size)
            d.d_pair_x, d.d_pair_y = base + o["x"], base + o["y"]
            d.d_lo, d.d_hi = wbase + o["lo"], wbase + o["hi"]
            d.d_xedges, d.d_yedges = sbase + o["xe"], sbase + o["ye"]
            d.d_hist, d.d_rank = wbase + o["hist"], wbase + o["rank"]
            d.d_rankpos, d.d_sort_tmp = wbase + o["rankpos"], wbase + o["tmp"]
        self.d_desc = _upload_structs(self.h_desc)
        self._stats = None

    def run(self, slopes_per_cta=0):
        rc = self.lib.rb2_hough_accumulate(self.n, self.d_desc.data_ptr(), C.addressof(self.h_desc),
                                           self.small.ptr(self.o_stats), int(slopes_per_cta), current_stream_ptr())
        N.check(rc, "rb2_hough_accumulate")
        self._stats = None
        return self

    # ---- results -----------------------------------------------------------
    @property
    def stats(self):
        if self._stats is None:
            raw = self.small.numpy(self.o_stats, np.uint8, C.sizeof(N.HoughStats) * self.n).tobytes()
            self._stats = (N.HoughStats * self.n).from_buffer_copy(raw)
            for s in self._stats:
                if s.status != 0:
                    N.check(s.status, "rb2_hough_accumulate (device status)")
        return self._stats

    def _dims(self, i):
        p = self.problems[i]
        return int(p["xbins"]), int(p["ybins"])

    def hist(self, i=0):
        xb, yb = self._dims(i)
        return self.work.numpy(self._off[i]["hist"], np.uint32, xb * yb).reshape(xb, yb)

    def xedges(self, i=0):
        return self.small.numpy(self._off[i]["xe"], np.float64, self._dims(i)[0] + 1)

    def yedges(self, i=0):
        return self.small.numpy(self._off[i]["ye"], np.float64, self._dims(i)[1] + 1)

    def rank(self, i=0):
        xb, yb = self._dims(i)
        return self.work.numpy(self._off[i]["rank"], np.int32, xb * yb)

    def intervals(self, i=0):
        P = self.h_desc[i].n_pairs
        return (self.work.numpy(self._off[i]["lo"], np.int32, P),
                self.work.numpy(self._off[i]["hi"], np.int32, P))

    def prefetch(self):
        """Bring every spectrum's edges + stats to the host in one copy (batch set-up)."""
        self.small.prefetch()
        self._stats = None

    def first_rows(self):
        """hist[:, 0] of every problem (the row the Hough-density weight needs),
        gathered on the device and copied once: list of float64 arrays."""
        torch = _torch()
        idx, lens = [], []
        for i in range(self.n):
            xb, yb = self._dims(i)
            idx.append(self._off[i]["hist"] // 4 + np.arange(xb, dtype=np.int64) * yb)
            lens.append(xb)
        flat = torch.from_numpy(np.concatenate(idx)).cuda()
        TRANSFER["h2d"] += flat.numel() * 8
        vals = self.work.buf.view(torch.int32)[flat].cpu().numpy().astype(np.float64)
        TRANSFER["d2h"] += vals.size * 4
        return np.split(vals, np.cumsum(lens)[:-1])

    def points(self, i=0):
        """HoughTransform.hough_points, materialised in the reference's row order."""
        torch = _torch()
        n_points = int(self.stats[i].n_points)
        S = self.h_desc[i].n_slopes
        off = torch.empty(S + 2, dtype=torch.int64, device="cuda")
        out = torch.empty((max(n_points, 1), 2), dtype=torch.float64, device="cuda")
        one = N.struct_array(N.HoughDesc, 1)
        C.memmove(one, C.byref(self.h_desc[i]), C.sizeof(N.HoughDesc))
        rc = self.lib.rb2_hough_points(C.addressof(one), off.data_ptr(), out.data_ptr(), n_points,
                                       current_stream_ptr())
        N.check(rc, "rb2_hough_points")
        return out[:n_points].cpu().numpy()


def candidate_points(hough, i, tol, gauss_denom, pairs=None):
    """Calibrator.candidates (calibrator.py:239-261) of problem i of a HoughBatch: (offsets[B + 1], x, y, w) -
    line r in rank order owns entries offsets[r]:offsets[r + 1] (rb2_candidate_points; on demand only).
    pairs: [P, 2] (pixel, wavelength) when the batch does not hold them (explicit-points transforms)."""
    torch = _torch()
    lib = N.load()
    one = N.struct_array(N.HoughDesc, 1)
    C.memmove(one, C.byref(hough.h_desc[i]), C.sizeof(N.HoughDesc))
    keep = None
    if pairs is not None:
        pairs = np.asarray(pairs, dtype=np.float64).reshape(-1, 2)
        keep = (torch.from_numpy(np.ascontiguousarray(pairs[:, 0])).cuda(), torch.from_numpy(np.ascontiguousarray(pairs[:, 1])).cuda())
        TRANSFER["h2d"] += pairs.nbytes
        one[0].d_pair_x, one[0].d_pair_y, one[0].n_pairs = keep[0].data_ptr(), keep[1].data_ptr(), len(pairs)
    B = one[0].xbins * one[0].ybins
    off = torch.zeros(B + 1, dtype=torch.int64, device="cuda")
    args = (C.addressof(one), float(tol), float(gauss_denom), off.data_ptr())
    N.check(lib.rb2_candidate_points(*args, None, None, None, current_stream_ptr()), "rb2_candidate_points")
    torch.cumsum(off, 0, out=off)
    total = int(off[B].item())
    out = torch.empty((3, max(total, 1)), dtype=torch.float64, device="cuda")
    if total:
        N.check(lib.rb2_candidate_points(*args, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                         current_stream_ptr()), "rb2_candidate_points")
    h = out[:, :total].cpu().numpy()
    TRANSFER["d2h"] += h.nbytes + 8 * (B + 1)
    del keep
    return off.cpu().numpy(), h[0], h[1], h[2]


class PointsHoughBatch(HoughBatch):
    """bin_hough_points for EXPLICIT Hough points (brute force / add_hough_points / load):
    rb2_hough_bin_points.  problems: dicts with points ([N, 2] gradient, intercept), xbins, ybins.
    Same result interface as HoughBatch (hist / edges / rank, consumed in place by VoteBatch)."""

    def __init__(self, problems):
        self.lib = N.load()
        _torch()
        self.n = len(problems)
        self.problems = problems
        ar, wk, sm = Arena(), Arena(), Arena()
        self.arena, self.work, self.small = ar, wk, sm
        self.h_desc = N.struct_array(N.HoughDesc, self.n)
        self._off = []
        for p in problems:
            pts = np.asarray(p["points"], dtype=np.float64).reshape(-1, 2)
            if not np.all(np.isfinite(pts)):  # np.histogram2d: "autodetected range of [nan, nan] is not finite"
                raise ValueError("hough_points must be finite")
            xb, yb = int(p["xbins"]), int(p["ybins"])
            B = xb * yb
            self._off.append(dict(x=ar.add(np.ascontiguousarray(pts[:, 0])), y=ar.add(np.ascontiguousarray(pts[:, 1])),
                                  n=len(pts), xe=sm.reserve(8 * (xb + 1)), ye=sm.reserve(8 * (yb + 1)),
                                  hist=wk.reserve(4 * B), rank=wk.reserve(4 * B), rankpos=wk.reserve(4 * B),
                                  tmp=wk.reserve(4 * B)))
        self.o_stats = sm.reserve(C.sizeof(N.HoughStats) * self.n)
        base, wbase, sbase = ar.upload(), wk.upload(), sm.upload()
        for i, p in enumerate(problems):
            o, d = self._off[i], self.h_desc[i]
            d.n_slopes, d.xbins, d.ybins, d.n_pairs = 0, int(p["xbins"]), int(p["ybins"]), o["n"]
            d.d_pair_x, d.d_pair_y = base + o["x"], base + o["y"]
            d.d_xedges, d.d_yedges = sbase + o["xe"], sbase + o["ye"]
            d.d_hist, d.d_rank = wbase + o["hist"], wbase + o["rank"]
            d.d_rankpos, d.d_sort_tmp = wbase + o["rankpos"], wbase + o["tmp"]
        self.d_desc = _upload_structs(self.h_desc)
        self._stats = None

    def run(self, slopes_per_cta=0):
        rc = self.lib.rb2_hough_bin_points(self.n, self.d_desc.data_ptr(), C.addressof(self.h_desc),
                                           self.small.ptr(self.o_stats), current_stream_ptr())
        N.check(rc, "rb2_hough_bin_points")
        self._stats = None
        return self

    def points(self, i=0):
        return np.asarray(self.problems[i]["points"], dtype=np.float64).reshape(-1, 2)


def brute_force_points(x, y, min_slope, max_slope, min_intercept, max_intercept):
    """HoughTransform.generate_hough_points_brute_force (houghtransform.py:94-140) on the device:
    [N, 2] (gradient, intercept) rows in the reference's order for a STABLE sort of x (np.argsort's
    default is not stable; the multiset of points does not depend on the tie order)."""
    torch = _torch()
    lib = N.load()
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    order = np.argsort(x, kind="stable")
    dx, dy = torch.from_numpy(x[order]).cuda(), torch.from_numpy(y[order]).cuda()
    TRANSFER["h2d"] += 16 * len(x)
    n = len(x)
    cnt = torch.zeros(max(n, 1), dtype=torch.int64, device="cuda")
    off = torch.zeros(max(n, 1) + 1, dtype=torch.int64, device="cuda")
    args = (dx.data_ptr(), dy.data_ptr(), n, float(min_slope), float(max_slope), float(min_intercept),
            float(max_intercept), cnt.data_ptr(), off.data_ptr())
    N.check(lib.rb2_hough_brute_force(*args, None, current_stream_ptr()), "rb2_hough_brute_force")
    total = int(off[max(n - 1, 0)].item())
    TRANSFER["d2h"] += 8
    pts = torch.empty((max(total, 1), 2), dtype=torch.float64, device="cuda")
    if total:
        N.check(lib.rb2_hough_brute_force(*args, pts.data_ptr(), current_stream_ptr()), "rb2_hough_brute_force")
    TRANSFER["d2h"] += 16 * total
    return pts[:total].cpu().numpy()


# ---------------------------------------------------------------------------
# K2  candidate vote
# ---------------------------------------------------------------------------


def vote_layout(pairs):
    """Host-side CSR layout of the pair list for K2.

    Returns (upeaks, pair_off, order, pair_y_csr, pair_wid, uwaves): unique
    peaks ascending, CSR offsets, the permutation pairs -> CSR (original order
    inside a peak), the wavelengths in CSR order and their ids among the sorted
    distinct wavelengths.
    """
    pairs = np.asarray(pairs, dtype=np.float64).reshape(-1, 2)
    upeaks, inv = np.unique(pairs[:, 0], return_inverse=True)
    order = np.argsort(inv, kind="stable")
    counts = np.bincount(inv, minlength=len(upeaks))
    pair_off = np.zeros(len(upeaks) + 1, dtype=np.int32)
    np.cumsum(counts, out=pair_off[1:])
    y = pairs[order, 1]
    uwaves, wid = np.unique(y, return_inverse=True)
    return upeaks, pair_off, order, y, wid.astype(np.int32), uwaves


def vote_layout_product(peaks, lines):
    """vote_layout for pairs = product(peaks, lines) (calibrator.py:103-106) without building
    the pair list, when the peaks are strictly increasing (else falls back to the general path)."""
    peaks = np.asarray(peaks, dtype=np.float64)
    lines = np.asarray(lines, dtype=np.float64)
    n_p, n_a = len(peaks), len(lines)
    if n_p > 1 and not np.all(np.diff(peaks) > 0):
        return vote_layout(np.column_stack((np.repeat(peaks, n_a), np.tile(lines, n_p))))
    uwaves, inv = np.unique(lines, return_inverse=True)
    pair_off = (np.arange(n_p + 1, dtype=np.int64) * n_a).astype(np.int32)
    return (peaks, pair_off, np.arange(n_p * n_a), np.tile(lines, n_p),
            np.tile(inv.astype(np.int32), n_p), uwaves)


class VoteBatch:
    """Per-peak candidate vote for a batch of spectra (rb2_candidate_vote).

    `hough` is the HoughBatch that produced the histogram (its device buffers
    are consumed in place); `settings[i]` = dict(pairs, top_n, weighted, tol,
    gauss_denom).
    """

    def __init__(self, hough: HoughBatch, settings, diagnostics=False):
        self.lib = N.load()
        self.n = hough.n
        assert len(settings) == self.n
        self.hough = hough
        ar, out = Arena(), Arena()
        self.arena, self.out = ar, out
        self.h_desc = N.struct_array(N.VoteDesc, self.n)
        self._off = []
        self.layouts = []
        for i, s in enumerate(settings):
            if s.get("pairs") is None:
                upeaks, pair_off, order, y, wid, uwaves = vote_layout_product(s["peaks"], s["lines"])
            else:
                upeaks, pair_off, order, y, wid, uwaves = vote_layout(s["pairs"])
            self.layouts.append((upeaks, pair_off, order, y, wid, uwaves))
            nu, nw, top_n = len(upeaks), len(uwaves), int(s["top_n"])
            o = dict(ux=ar.add(upeaks), off=ar.add(pair_off), y=ar.add(y), wid=ar.add(wid),
                     cw=out.reserve(8 * nu * top_n), cp=out.reserve(4 * nu * top_n), cn=out.reserve(4 * nu),
                     st=out.reserve(4))
            if diagnostics:
                o["agg"] = out.reserve(8 * nu * nw)
                o["cnt"] = out.reserve(4 * nu * nw)
            self._off.append(o)
        base, obase = ar.upload(), out.upload()
        out.buf.zero_()
        for i, s in enumerate(settings):
            o = self._off[i]
            upeaks, pair_off, order, y, wid, uwaves = self.layouts[i]
            d = self.h_desc[i]
            hd = hough.h_desc[i]
            d.n_upeaks, d.n_uwaves = len(upeaks), len(uwaves)
            d.xbins, d.ybins = hd.xbins, hd.ybins
            d.top_n, d.weighted = int(s["top_n"]), int(bool(s["weighted"]))
            d.tol, d.gauss_denom = float(s["tol"]), float(s["gauss_denom"])
            d.d_upeak_x, d.d_pair_off = base + o["ux"], base + o["off"]
            d.d_pair_y, d.d_pair_wid = base + o["y"], base + o["wid"]
            d.d_xedges, d.d_yedges, d.d_rankpos = hd.d_xedges, hd.d_yedges, hd.d_rankpos
            d.d_rank = hd.d_rank
            d.d_cand_wave, d.d_cand_pair, d.d_cand_n = obase + o["cw"], obase + o["cp"], obase + o["cn"]
            d.d_agg = obase + o["agg"] if diagnostics else None
            d.d_cnt = obase + o["cnt"] if diagnostics else None
            d.d_status = obase + o["st"]
        self.d_desc = _upload_structs(self.h_desc)

    def run(self):
        rc = self.lib.rb2_candidate_vote(self.n, self.d_desc.data_ptr(), C.addressof(self.h_desc),
                                         current_stream_ptr())
        N.check(rc, "rb2_candidate_vote")
        return self

    def candidates(self, i=0):
        """(candidate_peak, candidate_arc) as the reference's lists
        (calibrator.py:174-217): ascending unique peaks, n entries each."""
        o = self._off[i]
        d = self.h_desc[i]
        st = int(self.out.numpy(o["st"], np.int32, 1)[0])
        if st != 0:
            N.check(st, "rb2_candidate_vote (device status)")
        nu, top_n = d.n_upeaks, d.top_n
        cn = self.out.numpy(o["cn"], np.int32, nu)
        cw = self.out.numpy(o["cw"], np.float64, nu * top_n).reshape(nu, top_n)
        upeaks = self.layouts[i][0]
        keep = np.arange(top_n)[None, :] < cn[:, None]
        cand_peak = np.repeat(upeaks, cn)
        cand_arc = cw[keep]
        return cand_peak, cand_arc

    def aggregates(self, i=0):
        o = self._off[i]
        d = self.h_desc[i]
        return (self.out.numpy(o["agg"], np.float64, d.n_upeaks * d.n_uwaves).reshape(d.n_upeaks, d.n_uwaves),
                self.out.numpy(o["cnt"], np.int32, d.n_upeaks * d.n_uwaves).reshape(d.n_upeaks, d.n_uwaves))

    def prefetch(self):
        """Every spectrum's candidates to the host in one copy."""
        self.out.prefetch()


# ---------------------------------------------------------------------------
# a7  RANSAC setup (host; calibrator.py:441-523) -> K3 inputs
# ---------------------------------------------------------------------------


def sampling_cdf32(peaks):
    """The reference's sampling law (calibrator.py:521-523: inverse 10-bin density of the peaks,
    incl. the index -1 wrap-around of np.digitize(...) - 1) as (prob, uint32 CDF).  Every sum is
    sequential so that rb2_ransac_prepare reproduces the table bit for bit on the device."""
    lo, hi = peaks[0], peaks[-1]
    if lo == hi:
        lo, hi = lo - 0.5, hi + 0.5
    edges = np.linspace(lo, hi, 11)  # np.histogram(peaks, bins=10)[1]
    idx = np.minimum(np.searchsorted(edges, peaks, side="right") - 1, 9)
    counts = np.bincount(idx, minlength=10)
    dig = np.searchsorted(edges, peaks, side="left") - 1  # np.digitize(peaks, edges, right=True) - 1
    if not np.all(counts[dig] > 0):
        # a peak exactly on an interior histogram edge whose left bin is empty: the reference's weight is
        # 1/0 = inf, prob = nan, and np.random.choice raises this on the first draw (calibrator.py:538)
        raise ValueError("probabilities contain NaN")
    w = 1.0 / counts[dig]
    prob = w / np.cumsum(w)[-1]
    cdf = np.cumsum(prob)
    cdf /= cdf[-1]
    c32 = np.minimum(np.floor(cdf * 4294967296.0), 4294967295.0).astype(np.uint32)
    c32[-1] = 0xFFFFFFFF
    return prob, c32


def basis_tables(fit_type, deg):
    """fit_type 'legendre' / 'chebyshev' (calibrator.py:1625-1631) for K3: (code, c0_row[deg+1], grad[deg, deg+1]).
    c0_row: first row of the monomial -> basis change matrix (the intercept test reads the basis' first
    coefficient); grad: the monomial coefficients of sum_k (k+1) c_(k+1) B_k(x) - util._derivative applied to
    the basis coefficients, evaluated with the basis' polyval (:614-618) - as a linear map of the hypothesis'
    monomial coefficients."""
    code = N.BASIS_CODES[fit_type]
    if code == 0:
        return 0, None, None
    to_basis, to_poly = ((np.polynomial.legendre.poly2leg, np.polynomial.legendre.leg2poly) if code == 1 else
                         (np.polynomial.chebyshev.poly2cheb, np.polynomial.chebyshev.cheb2poly))
    n = deg + 1

    def col(fn, k, size):
        e = np.zeros(k + 1)
        e[k] = 1.0
        out = np.zeros(size)
        v = np.atleast_1d(fn(e))
        out[:len(v)] = v
        return out

    T = np.column_stack([col(to_basis, j, n) for j in range(n)])         # c_basis = T @ a_mono
    D = np.zeros((deg, n))
    D[np.arange(deg), np.arange(1, n)] = np.arange(1, n)                  # d_k = (k + 1) c_(k+1)
    Bk = np.column_stack([col(to_poly, k, deg) for k in range(deg)])      # monomials of sum_k d_k B_k
    return code, np.ascontiguousarray(T[0]), np.ascontiguousarray(Bk @ D @ T)


def bspline_basis_pp(t):
    """The cubic B-spline basis on knots t (first / last four equal) as local polynomials: for every knot
    interval m = 0..n-4 (n = len(t) - 4 basis functions) the four non-zero basis functions N_m..N_m+3 in
    powers of (x - t[m+3]), ascending: array [n-3, 4, 4], flattened."""
    from scipy.interpolate import BSpline

    t = np.asarray(t, dtype=np.float64)
    n = len(t) - 4
    b = BSpline(t, np.eye(n), 3)
    left = t[3:n]  # left end of interval m is t[m + 3] (evaluation there is right-continuous)
    fact = (1.0, 1.0, 2.0, 6.0)
    tay = np.stack([b(left, nu=p) / fact[p] for p in range(4)], axis=2)  # [interval, basis, power]
    m = np.arange(n - 3)
    out = np.stack([tay[m, m + a, :] for a in range(4)], axis=1)  # [interval, 4 non-zero basis functions, power]
    return np.ascontiguousarray(out.reshape(-1))


class RansacSetup:
    """Read-only inputs of the hypothesis loop, as the reference prepares them
    at the top of _solve_candidate_ransac (calibrator.py:441-523)."""

    def __init__(self, cand_peak, cand_arc, *, fit_deg, sample_size, ransac_tolerance, min_intercept,
                 max_intercept, pixel_list, hist, xedges, yedges, hough_weight=1.0, filter_close=False,
                 minimum_matches=0, minimum_peak_utilisation=0.0, minimum_fit_error=1e-4, n_peaks_total=None,
                 pix_known=None, wave_known=None, cost_ceiling=1e50, fit_type="poly"):
        self.fit_type = fit_type
        self.basis, self.basis_c0, self.basis_grad = basis_tables(fit_type, int(fit_deg))
        x = np.array(cand_peak, dtype=np.float64)
        y = np.array(cand_arc, dtype=np.float64)
        if filter_close:  # calibrator.py:457-464: drop the lower wavelength of every close pair
            uy = np.unique(y)
            close = uy[:-1][(uy[1:] - uy[:-1]) < 3 * ransac_tolerance]
            keep = ~np.isin(y, close)
            x, y = x[keep], y[keep]
        self.x, self.y = x, y
        self.deg = int(fit_deg)
        self.sample_size = int(sample_size)
        self.tol = float(ransac_tolerance)
        self.min_intercept, self.max_intercept = float(min_intercept), float(max_intercept)
        self.hough_weight = float(hough_weight)
        self.minimum_matches = int(minimum_matches)
        self.minimum_peak_utilisation = float(minimum_peak_utilisation)
        self.minimum_fit_error = float(minimum_fit_error)
        self.cost_ceiling = float(cost_ceiling)
        self.peaks = np.unique(x)  # sorted
        self.n_peaks_total = len(self.peaks) if n_peaks_total is None else int(n_peaks_total)
        self.enough = len(self.peaks) > self.deg  # calibrator.py:468
        self.known_x = np.zeros(0) if pix_known is None else np.asarray(pix_known, dtype=np.float64).reshape(-1)
        self.known_y = np.zeros(0) if wave_known is None else np.asarray(wave_known, dtype=np.float64).reshape(-1)
        # candidate table: CSR over ascending peaks, original order inside a peak (:494-495)
        order = np.argsort(x, kind="stable")
        self.cand_wave = y[order]
        counts = np.searchsorted(x[order], self.peaks, side="right") - np.searchsorted(x[order], self.peaks, side="left")
        self.cand_off = np.zeros(len(self.peaks) + 1, dtype=np.int32)
        np.cumsum(counts, out=self.cand_off[1:])
        self.uwaves, wid = np.unique(self.cand_wave, return_inverse=True)
        self.cand_wid = wid.astype(np.int32)
        # sampling law (:521-523), index -1 wrap-around kept
        if self.enough:
            self.prob, self.cdf32 = sampling_cdf32(self.peaks)
        else:
            self.prob = None
            self.cdf32 = np.zeros(len(self.peaks), dtype=np.uint32)
        ptp = np.ptp(self.peaks) if len(self.peaks) else 0.0
        self.pix_min = float(self.peaks[0] - ptp * 0.2) if len(self.peaks) else 0.0  # :585
        self.pix_max = float(self.peaks[-1] + ptp * 0.2) if len(self.peaks) else 0.0  # :586
        # pixel list: NULL on the device when it is arange(n)
        pl = np.asarray(pixel_list, dtype=np.float64)
        self.n_pix = len(pl)
        self.pixel_list = pl
        self.pix_is_arange = bool(len(pl) and np.array_equal(pl, np.arange(len(pl), dtype=np.float64)))
        self.pix_sorted = bool(np.all(np.diff(pl) > 0)) if len(pl) > 1 else True
        # Hough-density weight (calibrator.py:497-506, 614-627)
        self._weight_tables(np.asarray(hist, dtype=np.float64), np.asarray(xedges), np.asarray(yedges))

    def _weight_tables(self, hist, xedges, yedges):
        from scipy import interpolate

        hx = (xedges[1] - xedges[0]) / 2.0
        hy = (yedges[1] - yedges[0]) / 2.0
        gc = xedges[1:] - hx
        ic = yedges[1:] - hy
        self.gc_min, self.gc_max, self.ic_min = float(gc[0]), float(gc[-1]), float(ic[0])
        if not np.isfinite(self.hough_weight):
            raise NotImplementedError("hough_weight must be finite")
        general = bool(self.basis or hist.size <= 65536 or self.ic_min <= 20.0 * max(abs(self.gc_min), abs(self.gc_max)))
        self.bs_c = None
        if not general:
            # fine binnings (config 5: 10^6 bins) on an Angstrom scale: only the section of the spline along the
            # first intercept row is ever evaluated - the interpolating cubic spline of
            # hist[:, 0] (FITPACK's knots are the not-a-knot ones), without fitting the 2-D spline (0.15 s)
            b = interpolate.make_interp_spline(gc, hist[:, 0], k=3)
            brk = np.unique(b.t)
            left = brk[:-1]
            fact = (1.0, 1.0, 2.0, 6.0)
            coef = np.stack([b(left, nu=k) / fact[k] for k in range(4)], axis=1)
            self.h_lo, self.h_hi = float(hist[0, 0]), float(hist[-1, 0])
            self.pp_breaks = brk.astype(np.float64)
            self.pp_coef = np.ascontiguousarray(coef.reshape(-1))
            return
        # the reference's own spline object gives the two clamped corner values and the
        # 1-D section along the first intercept row (evaluated as spline(t, ic_min))
        spl = interpolate.RectBivariateSpline(gc, ic, hist)
        self.h_lo = float(spl(gc[0], ic[0], grid=False))
        self.h_hi = float(spl(gc[-1], ic[0], grid=False))
        # pp-form of t -> spline(t, ic_min) between the FITPACK knots: a cubic on each knot
        # interval, recovered exactly from 4 samples per interval
        tx = spl.get_knots()[0]
        brk = np.unique(tx)
        npp = len(brk) - 1
        u = np.array([0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0])
        a, w = brk[:-1], np.diff(brk)
        v = spl((a[:, None] + u[None, :] * w[:, None]).ravel(), np.full(4 * npp, ic[0]), grid=False).reshape(npp, 4)
        coef = np.linalg.solve(np.vander(u, 4, increasing=True), v.T).T / w[:, None] ** np.arange(4)[None, :]
        self.pp_breaks = brk.astype(np.float64)
        self.pp_coef = coef.reshape(-1)
        # the spline itself, for hypotheses outside the corner regime (general evaluation on the device): nm-scale
        # problems, a curvy degree-5 hypothesis now and then on an Angstrom scale too (242 of 490 000 scored ones on
        # the GMOS arc), and every hypothesis of a non-monomial basis (rb2_ransac_desc.basis)
        ty = spl.get_knots()[1]
        self.bs_tx, self.bs_ty = np.ascontiguousarray(tx, dtype=np.float64), np.ascontiguousarray(ty, dtype=np.float64)
        self.bs_c = np.ascontiguousarray(spl.get_coeffs(), dtype=np.float64)  # [(len(tx)-4) * (len(ty)-4)], x first
        self.bs_ppx, self.bs_ppy = bspline_basis_pp(self.bs_tx), bspline_basis_pp(self.bs_ty)

    def prior_cost(self, fit_coeff):
        """The initial cost of a pre-existing fit (calibrator.py:512-515): the plain sum of the
        bijective-match errors of `fit_coeff` (unweighted, unclipped) - the ceiling a hypothesis has to
        beat.  Part of the loop's set-up, evaluated once on the host."""
        c = np.asarray(fit_coeff, dtype=np.float64)
        polyval = {0: np.polynomial.polynomial.polyval, 1: np.polynomial.legendre.legval,
                   2: np.polynomial.chebyshev.chebval}[self.basis]  # self.polyval of the fit (:1621-1631)
        fit = polyval(self.peaks, c)
        best_e, best_w = np.empty(len(self.peaks)), np.empty(len(self.peaks))
        for k in range(len(self.peaks)):
            cand = self.cand_wave[self.cand_off[k]:self.cand_off[k + 1]]
            e = np.abs(fit[k] - cand)
            i = int(np.argmin(e))  # first minimum (:349-351)
            best_e[k], best_w[k] = e[i], cand[i]
        err = [np.min(best_e[best_w == w]) for w in np.unique(best_w)]  # one peak per wavelength (:356-376)
        return float(sum(err))

    @property
    def min_inliers(self):
        # inliers > deg (calibrator.py:645), inliers >= minimum_matches (:684)
        return max(self.deg + 1, self.minimum_matches)

    @property
    def min_inliers_frac(self):
        return self.minimum_peak_utilisation * self.n_peaks_total  # :692-700


def weight_upper_bound(s):
    """Rigorous upper bound of the Hough-density weight of ANY hypothesis of problem `s`
    (every pixel contributes h_lo, h_hi or a value of the section spline): the Bernstein
    coefficients of each cubic piece enclose its range.  <= 0 disables K3's early skip (a
    spline that can go negative makes the cost unbounded below)."""
    c = np.asarray(s.pp_coef, dtype=np.float64).reshape(-1, 4)
    h = np.diff(np.asarray(s.pp_breaks, dtype=np.float64))
    b0 = c[:, 0]
    b1 = c[:, 0] + c[:, 1] * h / 3.0
    b2 = c[:, 0] + 2.0 * c[:, 1] * h / 3.0 + c[:, 2] * h * h / 3.0
    b3 = c[:, 0] + c[:, 1] * h + c[:, 2] * h * h + c[:, 3] * h ** 3
    bern = np.stack((b0, b1, b2, b3))
    lo = min(float(bern.min()), s.h_lo, s.h_hi)
    hi = max(float(bern.max()), s.h_lo, s.h_hi)
    if not (lo >= 0.0 and hi > 0.0 and s.hough_weight > 0.0 and np.isfinite(hi)):
        return -1.0
    return float(s.hough_weight * s.n_pix * hi * (1.0 + 1e-9))


_SCRATCH = {}


def ransac_scratch(nbytes):
    """K3's survivor queues (rb2_ransac_scratch_bytes: 18 GB at the default chunk capacity of 2^25 hypotheses and three lanes):
    ONE grow-only buffer per device and process.  Every K3 call is ordered on the caller's stream, so calls
    can share it; allocating it per call would cost more than most fits."""
    torch = _torch()
    dev = torch.cuda.current_device()
    t = _SCRATCH.get(dev)
    if t is None or t.numel() < nbytes:
        _SCRATCH[dev] = None
        t = _SCRATCH[dev] = torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device="cuda")
    return t


class RansacBatch:
    """K3 + K4 for a batch of prepared problems (rb2_ransac_batch / rb2_refit_winners)."""

    def __init__(self, setups, blocks_per_problem=0):
        self.lib = N.load()
        _torch()
        self.setups = list(setups)
        self.n = len(self.setups)
        ar = Arena()
        self.arena = ar
        self.h_desc = N.struct_array(N.RansacDesc, self.n)
        self._off = []
        for s in self.setups:
            o = dict(peaks=ar.add(s.peaks), coff=ar.add(s.cand_off), cwave=ar.add(s.cand_wave),
                     cwid=ar.add(s.cand_wid), cdf=ar.add(s.cdf32), kx=ar.add(s.known_x), ky=ar.add(s.known_y),
                     ppb=ar.add(s.pp_breaks), ppc=ar.add(s.pp_coef))
            if getattr(s, "basis", 0):
                o.update(bc0=ar.add(s.basis_c0), bgr=ar.add(s.basis_grad))
            if getattr(s, "bs_c", None) is not None:
                o.update(btx=ar.add(s.bs_tx), bty=ar.add(s.bs_ty), bc=ar.add(s.bs_c), bpx=ar.add(s.bs_ppx),
                         bpy=ar.add(s.bs_ppy))
            if not s.pix_is_arange:
                if not s.pix_sorted:
                    raise NotImplementedError("pixel_list must be strictly increasing")
                o["pix"] = ar.add(s.pixel_list)
            self._off.append(o)
        self.max_up = max((len(s.peaks) for s in self.setups), default=1)
        base = ar.upload()
        for i, s in enumerate(self.setups):
            o = self._off[i]
            d = self.h_desc[i]
            d.n_upeaks, d.n_wids, d.n_cand = len(s.peaks), len(s.uwaves), len(s.cand_wave)
            d.d_peaks, d.d_cand_off = base + o["peaks"], base + o["coff"]
            d.d_cand_wave, d.d_cand_wid, d.d_cdf32 = base + o["cwave"], base + o["cwid"], base + o["cdf"]
            d.n_known = len(s.known_x)
            d.d_known_x, d.d_known_y = base + o["kx"], base + o["ky"]
            d.deg, d.sample_size = s.deg, s.sample_size
            d.min_intercept, d.max_intercept = s.min_intercept, s.max_intercept
            d.pix_min, d.pix_max, d.tol = s.pix_min, s.pix_max, s.tol
            d.hough_weight, d.gc_min, d.gc_max, d.ic_min = s.hough_weight, s.gc_min, s.gc_max, s.ic_min
            d.h_lo, d.h_hi = s.h_lo, s.h_hi
            d.n_pp = len(s.pp_breaks) - 1
            d.d_pp_breaks, d.d_pp_coef = base + o["ppb"], base + o["ppc"]
            d.n_pix = s.n_pix
            d.w_upper = weight_upper_bound(s)
            d.d_pix = None if s.pix_is_arange else base + o["pix"]
            d.min_inliers, d.min_inliers_frac = s.min_inliers, s.min_inliers_frac
            d.cost_ceiling = s.cost_ceiling
            d.floor_cost, d.floor_id = 0.0, -1
            d.min_fit_error = float(s.minimum_fit_error)
            if "bc0" in o:
                d.basis, d.d_basis_c0, d.d_basis_grad = int(s.basis), base + o["bc0"], base + o["bgr"]
            if "bc" in o:
                d.bs_nx, d.bs_ny = len(s.bs_tx) - 4, len(s.bs_ty) - 4
                d.d_bs_tx, d.d_bs_ty, d.d_bs_c = base + o["btx"], base + o["bty"], base + o["bc"]
                d.d_bs_ppx, d.d_bs_ppy = base + o["bpx"], base + o["bpy"]
        torch = _torch()
        sms = C.c_int(0)
        N.check(self.lib.rb2_sm_count(C.byref(sms)), "rb2_sm_count")
        self.sm_count = sms.value
        self.blocks_per_problem = max(0, int(blocks_per_problem))  # stage-A CTAs per problem per chunk; 0 = auto
        self.scratch = ransac_scratch(self.lib.rb2_ransac_scratch_bytes(self.n, self.blocks_per_problem))
        # every per-fit output lives in ONE device buffer so that a fit costs one D2H copy:
        # [winner records | refit results | matched_x | matched_y | residual]
        rsz, fsz = C.sizeof(N.RansacResult) * self.n, C.sizeof(N.RefitResult) * self.n
        msz = 8 * self.n * self.max_up
        self._o_res, self._o_fit, self._o_mx = 0, rsz, rsz + fsz
        self._o_my, self._o_rs = self._o_mx + msz, self._o_mx + 2 * msz
        self.d_out = torch.empty(rsz + fsz + 3 * msz, dtype=torch.uint8, device="cuda")
        self.h_out = torch.empty(rsz + fsz + 3 * msz, dtype=torch.uint8, pin_memory=True)
        self.d_result = self.d_out[:rsz]
        self.d_refit = self.d_out[rsz:rsz + fsz]
        self.d_desc = torch.empty(C.sizeof(N.RansacDesc) * self.n, dtype=torch.uint8, device="cuda")
        self.h_desc_pinned = torch.empty(C.sizeof(N.RansacDesc) * self.n, dtype=torch.uint8, pin_memory=True)
        self._desc_synced = False
        self._desc_event = None
        self.d_gather = None
        self._keep = []

    def _sync_desc(self):
        """Descriptors -> device (pinned staging, async on the current stream)."""
        torch = _torch()
        nb = C.sizeof(N.RansacDesc) * self.n
        if self._desc_event is not None:
            self._desc_event.synchronize()  # the previous async copy has read the staging buffer
        C.memmove(self.h_desc_pinned.data_ptr(), C.addressof(self.h_desc), nb)
        self.d_desc.copy_(self.h_desc_pinned, non_blocking=True)
        self._desc_event = torch.cuda.Event()
        self._desc_event.record()
        TRANSFER["h2d"] += nb
        self._desc_synced = True

    def run(self, hyp_begin, n_hyp, seed=0, replay=None, per_hypothesis=False, floors=None, replay_coeffs=None,
            samples=False, status_only=False, device_outputs=False):
        """Evaluate hypothesis ids [hyp_begin, hyp_begin + n_hyp) of every problem; the winner
        records stay on the device (results() / fetch() bring them back).

        replay: optional (x_idx[n_hyp, s], y_idx[n_hyp, s]) int arrays (parity
        mode, one problem), replay_coeffs: with it, [n_hyp, deg+1] hypothesis coefficients used
        instead of fitting (the reference's own polyfit output); per_hypothesis=True also returns
        status / cost / weight / coefficients / counts of every hypothesis (one problem);
        samples=True returns the draws themselves, "sample" [n_hyp, s, 2] = (peak index, candidate
        index inside that peak's list), -1 where the draw was aborted (one problem); status_only=True
        keeps the rb2_hyp_status byte of every hypothesis on the device (self.d_status, one problem).
        """
        torch = _torch()
        self._keep = []
        outs = None
        for i, s in enumerate(self.setups):
            d = self.h_desc[i]
            d.hyp_begin, d.n_hyp, d.seed = int(hyp_begin), int(n_hyp), int(seed)
            d.d_replay_x = d.d_replay_y = None
            d.d_out_status = d.d_out_cost = d.d_out_weight = d.d_out_coeffs = d.d_out_counts = None
            d.d_out_sample = d.d_replay_coeffs = None
            if floors is not None:
                d.floor_cost, d.floor_id = float(floors[i][0]), int(floors[i][1])
            else:
                d.floor_cost, d.floor_id = 0.0, -1
        if replay is not None:
            assert self.n == 1
            rx = torch.from_numpy(np.ascontiguousarray(replay[0], dtype=np.int32)).cuda()
            ry = torch.from_numpy(np.ascontiguousarray(replay[1], dtype=np.int32)).cuda()
            assert rx.shape == (n_hyp, self.setups[0].sample_size)
            self._keep += [rx, ry]
            self.h_desc[0].d_replay_x, self.h_desc[0].d_replay_y = rx.data_ptr(), ry.data_ptr()
            if replay_coeffs is not None:
                rc_ = torch.from_numpy(np.ascontiguousarray(replay_coeffs, dtype=np.float64)).cuda()
                assert rc_.shape == (n_hyp, self.setups[0].deg + 1)
                self._keep.append(rc_)
                self.h_desc[0].d_replay_coeffs = rc_.data_ptr()
        if per_hypothesis:
            assert self.n == 1
            nc = self.setups[0].deg + 1
            # per_hypothesis may name the fields wanted (the rest is not allocated nor written)
            want = ("status", "cost", "weight", "coeffs", "counts") if per_hypothesis is True else tuple(per_hypothesis)
            make = dict(status=lambda: torch.full((n_hyp,), 255, dtype=torch.uint8, device="cuda"),
                        cost=lambda: torch.full((n_hyp,), float("nan"), dtype=torch.float64, device="cuda"),
                        weight=lambda: torch.full((n_hyp,), float("nan"), dtype=torch.float64, device="cuda"),
                        coeffs=lambda: torch.full((n_hyp, nc), float("nan"), dtype=torch.float64, device="cuda"),
                        counts=lambda: torch.zeros((n_hyp, 2), dtype=torch.int32, device="cuda"))
            outs = {k: make[k]() for k in want}
            d = self.h_desc[0]
            ptr = lambda k: outs[k].data_ptr() if k in outs else None  # noqa: E731
            d.d_out_status, d.d_out_cost = ptr("status"), ptr("cost")
            d.d_out_weight, d.d_out_coeffs = ptr("weight"), ptr("coeffs")
            d.d_out_counts = ptr("counts")
        self.d_status = None
        if status_only:
            assert self.n == 1 and not per_hypothesis
            self.d_status = torch.full((n_hyp,), 255, dtype=torch.uint8, device="cuda")
            self.h_desc[0].d_out_status = self.d_status.data_ptr()
        if samples:
            assert self.n == 1
            outs = outs if outs is not None else {}
            outs["sample"] = torch.full((n_hyp, self.setups[0].sample_size, 2), -2, dtype=torch.int32, device="cuda")
            self.h_desc[0].d_out_sample = outs["sample"].data_ptr()
        self._sync_desc()
        rc = self.lib.rb2_ransac_batch(self.n, self.d_desc.data_ptr(), C.addressof(self.h_desc),
                                       self.d_result.data_ptr(), self.scratch.data_ptr(), self.blocks_per_problem,
                                       current_stream_ptr())
        N.check(rc, "rb2_ransac_batch")
        if outs is not None and not device_outputs:  # device_outputs: the caller post-processes them on the device
            outs = {k: v.cpu().numpy() for k, v in outs.items()}
        return outs

    def exchange(self, group=None):
        """The ONE collective of a sharded fit: all-gather the per-rank winner records (NCCL,
        device to device) and merge them on the device (rb2_merge_winners).  No-op when the
        process group has a single rank."""
        import torch.distributed as dist

        torch = _torch()
        if group is False or not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        if self.d_gather is None or self.d_gather.numel() != world * self.d_result.numel():
            self.d_gather = torch.empty(world * self.d_result.numel(), dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(self.d_gather, self.d_result, group=group)
        rc = self.lib.rb2_merge_winners(self.n, world, self.d_gather.data_ptr(), self.d_result.data_ptr(),
                                        current_stream_ptr())
        N.check(rc, "rb2_merge_winners")

    def set_winners(self, winners):
        """Overwrite the device winner records with host structs (fall-through / tests)."""
        self.d_result.copy_(_upload_structs(winners))

    def results(self):
        return _download_structs(self.d_result, N.RansacResult, self.n)

    def refit_async(self, mode=0):
        """K4 on the winner records currently on the device; nothing is copied back."""
        mfe = float("nan")  # every problem's own minimum_fit_error (rb2_ransac_desc.min_fit_error)
        if not self._desc_synced:
            self._sync_desc()
        base = self.d_out.data_ptr()
        rc = self.lib.rb2_refit_winners(self.n, self.d_desc.data_ptr(), C.addressof(self.h_desc),
                                        self.d_result.data_ptr(), float(mfe), int(mode), self.d_refit.data_ptr(),
                                        base + self._o_mx, base + self._o_my, base + self._o_rs,
                                        self.max_up, current_stream_ptr())
        N.check(rc, "rb2_refit_winners")

    def fetch_async(self):
        """Enqueue the ONE device->host copy of everything a fit produced on the current stream and return the
        event that marks its end (fetch_wait(event) reads the copy).  Lets a caller that runs fits back to back
        put K4 + this copy of fit k on a side stream while K3 of fit k + 1 runs (bench.py)."""
        torch = _torch()
        self.h_out.copy_(self.d_out, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        TRANSFER["d2h"] += self.h_out.numel()
        return ev

    def fetch_wait(self, event=None):
        """(winner records, refit results, matched_x, matched_y, residual) from the pinned host copy."""
        if event is not None:
            event.synchronize()
        raw = self.h_out.numpy()
        win = (N.RansacResult * self.n).from_buffer_copy(raw[self._o_res:self._o_fit].tobytes())
        fit = (N.RefitResult * self.n).from_buffer_copy(raw[self._o_fit:self._o_mx].tobytes())
        shp = (self.n, self.max_up)
        mx = raw[self._o_mx:self._o_my].view(np.float64).reshape(shp).copy()
        my = raw[self._o_my:self._o_rs].view(np.float64).reshape(shp).copy()
        rs = raw[self._o_rs:].view(np.float64).reshape(shp).copy()
        return win, fit, mx, my, rs

    def fetch(self):
        """ONE device->host copy of everything a fit produced:
        (winner records, refit results, matched_x, matched_y, residual)."""
        return self.fetch_wait(self.fetch_async())

    def refit(self, winners=None, mode=0):
        """K4 on the current winners (or on explicitly supplied RansacResult structs)."""
        if winners is not None:
            self.set_winners(winners)
        self.refit_async(mode)
        _, fit, mx, my, rs = self.fetch()
        return fit, mx, my, rs


def counters_dict(res):
    return {k: int(res.counters[i]) for i, k in enumerate(N.COUNTER_NAMES)}


def robust_polyfit(x, y, deg, mode=0):
    """models.robust_polyfit (models.py:70-123) on the device for one point set or a list of
    point sets; returns coefficients (low order first) [, ...] plus (rms, nfev, status)."""
    torch = _torch()
    lib = N.load()
    single = np.ndim(x[0]) == 0
    xs = [np.asarray(x, dtype=np.float64)] if single else [np.asarray(v, dtype=np.float64) for v in x]
    ys = [np.asarray(y, dtype=np.float64)] if single else [np.asarray(v, dtype=np.float64) for v in y]
    off = np.zeros(len(xs) + 1, dtype=np.int32)
    np.cumsum([len(v) for v in xs], out=off[1:])
    ar = Arena()
    ox, oy, oo = ar.add(np.concatenate(xs)), ar.add(np.concatenate(ys)), ar.add(off)
    base = ar.upload()
    out = torch.empty(C.sizeof(N.RefitResult) * len(xs), dtype=torch.uint8, device="cuda")
    rc = lib.rb2_robust_polyfit(len(xs), base + oo, off.ctypes.data, base + ox, base + oy, int(deg), int(mode),
                                out.data_ptr(), current_stream_ptr())
    N.check(rc, "rb2_robust_polyfit")
    res = _download_structs(out, N.RefitResult, len(xs))
    outs = [(np.array(r.coeffs[:deg + 1]) if r.accepted else np.full(deg + 1, np.nan), r.rms, r.huber_iters, r.pad)
            for r in res]
    return outs[0] if single else outs


class AdjustPolyfit:
    """Calibrator._adjust_polyfit (calibrator.py:754-820) on the device (rb2_adjust_polyfit): the objective of
    match_peaks(refine=True).  Peaks, atlas lines and the sorted pixel list are uploaded once; __call__ evaluates
    one trial shift (what scipy.optimize.minimize asks for), many() a whole array of them."""

    def __init__(self, peaks, lines, pixel_list, fit, tolerance, min_frac, fit_type="poly"):
        torch = _torch()
        self.lib = N.load()
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()  # noqa: E731
        self.peaks, self.lines = up(np.asarray(peaks).reshape(-1)), up(np.asarray(lines).reshape(-1))
        self.pix, self.fit = up(np.sort(np.asarray(pixel_list, dtype=np.float64))), up(np.asarray(fit).reshape(-1))
        TRANSFER["h2d"] += 8 * (self.peaks.numel() + self.lines.numel() + self.pix.numel() + self.fit.numel())
        self.tolerance, self.min_frac, self.basis = float(tolerance), float(min_frac), int(N.BASIS_CODES[fit_type])
        self.n_evals = 0

    def many(self, deltas):
        torch = _torch()
        d = np.ascontiguousarray(np.atleast_2d(np.asarray(deltas, dtype=np.float64)))
        dd = torch.from_numpy(d).cuda()
        out = torch.empty(len(d), dtype=torch.float64, device="cuda")
        rc = self.lib.rb2_adjust_polyfit(len(d), dd.data_ptr(), d.shape[1], self.fit.data_ptr(), self.fit.numel(), self.basis,
                                         self.peaks.data_ptr(), self.peaks.numel(), self.lines.data_ptr(), self.lines.numel(),
                                         self.pix.data_ptr(), self.pix.numel(), self.tolerance, self.min_frac, out.data_ptr(),
                                         current_stream_ptr())
        N.check(rc, "rb2_adjust_polyfit")
        TRANSFER["h2d"] += d.nbytes
        TRANSFER["d2h"] += 8 * len(d)
        self.n_evals += len(d)
        return out.cpu().numpy()

    def __call__(self, delta):
        return float(self.many(np.asarray(delta, dtype=np.float64).reshape(1, -1))[0])


class MatchBatch:
    """K5: Calibrator.match_peaks (calibrator.py:1824-1956, refine=False) for a batch of
    solutions (rb2_match_peaks).

    problems: list of dicts with peaks, lines, tolerance, fit_deg, robust_refit and either
    coeffs (host array) or d_coeffs (device address of the coefficients, e.g. a K4 refit
    record still on the device) + n_coef."""

    def __init__(self, problems):
        self.lib = N.load()
        self.n = len(problems)
        self.problems = problems
        self.max_p = max(1, max(len(p["peaks"]) for p in problems))
        ar = Arena()
        offs = [(ar.add(np.asarray(p["peaks"], dtype=np.float64)), ar.add(np.asarray(p["lines"], dtype=np.float64)))
                for p in problems]
        self.h_desc = N.struct_array(N.MatchDesc, self.n)
        o_desc = ar.reserve(C.sizeof(N.MatchDesc) * self.n)
        self._o_res = ar.reserve(C.sizeof(N.MatchResult) * self.n)
        self._o_mx = ar.reserve(8 * self.n * self.max_p)
        self._o_my = ar.reserve(8 * self.n * self.max_p)
        self._o_rs = ar.reserve(8 * self.n * self.max_p)
        self._end = ar.size
        base = ar.upload()
        for i, p in enumerate(problems):
            d = self.h_desc[i]
            d.d_peaks, d.d_atlas = base + offs[i][0], base + offs[i][1]
            d.n_peaks, d.n_atlas = len(p["peaks"]), len(p["lines"])
            if p.get("d_coeffs"):
                d.d_coeffs, d.n_coef = int(p["d_coeffs"]), int(p["n_coef"])
            else:
                c = np.asarray(p["coeffs"], dtype=np.float64)
                d.d_coeffs, d.n_coef = None, len(c)
                for k, v in enumerate(c):
                    d.coeffs[k] = float(v)
            d.fit_deg = int(p["fit_deg"]) if p.get("fit_deg") is not None else d.n_coef - 1
            d.tolerance = float(p["tolerance"])
            d.robust_refit = int(bool(p.get("robust_refit", True)))
            d.basis = int(N.BASIS_CODES[p.get("fit_type", "poly")])
        torch = _torch()
        self.arena = ar
        self._desc_host = torch.empty(C.sizeof(N.MatchDesc) * self.n, dtype=torch.uint8, pin_memory=True)
        C.memmove(self._desc_host.data_ptr(), C.addressof(self.h_desc), self._desc_host.numel())
        ar.buf[o_desc:o_desc + self._desc_host.numel()].copy_(self._desc_host, non_blocking=True)
        TRANSFER["h2d"] += self._desc_host.numel()
        self._d_desc = base + o_desc
        self._base = base

    def run(self, mode=0):
        b = self._base
        rc = self.lib.rb2_match_peaks(self.n, self._d_desc, C.addressof(self.h_desc), int(mode), b + self._o_res,
                                      b + self._o_mx, b + self._o_my, b + self._o_rs, self.max_p,
                                      current_stream_ptr())
        N.check(rc, "rb2_match_peaks")
        return self

    def fetch(self):
        """ONE device->host copy: (match results, matched_x, matched_y, residual)."""
        raw = self.arena.buf[self._o_res:self._end].cpu().numpy()
        TRANSFER["d2h"] += raw.nbytes
        o = self._o_res
        res = (N.MatchResult * self.n).from_buffer_copy(raw[:C.sizeof(N.MatchResult) * self.n].tobytes())
        shp, nb = (self.n, self.max_p), 8 * self.n * self.max_p
        mx = raw[self._o_mx - o:self._o_mx - o + nb].view(np.float64).reshape(shp)
        my = raw[self._o_my - o:self._o_my - o + nb].view(np.float64).reshape(shp)
        rs = raw[self._o_rs - o:self._o_rs - o + nb].view(np.float64).reshape(shp)
        return res, mx, my, rs
