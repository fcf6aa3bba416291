"""The step before the calibration path, on the device.

Mirrors the two calls every rascal example and test makes to turn an arc spectrum
into the ``peaks`` it hands to ``Calibrator`` (test/test_xe_calibration.py:69-70,
examples/example_*.py)::

    peaks, _ = scipy.signal.find_peaks(spectrum, height=..., prominence=..., distance=...)
    peaks = rascal.util.refine_peaks(spectrum, peaks, window_width=...)

* :func:`refine_peaks` - drop-in for ``rascal.util.refine_peaks`` (util.py:450-523):
  same arguments, same quirks (the in-place window normalisation, the start values,
  ``peak - window_width + centre``), same MINPACK Levenberg-Marquardt iteration as
  ``scipy.optimize.curve_fit`` runs for it.
* :func:`find_peaks` - ``scipy.signal.find_peaks`` for the conditions the reference's
  recipes use (``height``, ``distance``, ``prominence`` as scalar minima).
* :class:`PeakBatch` - both for a ragged batch of spectra with no host round trip in
  between; :func:`find_and_refine` is the one-call form.

Everything runs in hand-written CUDA kernels through the C ABI (rb2_find_peaks,
rb2_refine_peaks, include/rascal_b200.h); there is no CPU fallback.
"""
from __future__ import annotations


import numpy as np

from . import _native as N
from .engine import TRANSFER, _torch, current_stream_ptr

_UNSUPPORTED = ("threshold", "width", "wlen", "rel_height", "plateau_size")


class _Staging:
    """Grow-only pinned staging buffers, one per purpose, reused across calls (cudaHostAlloc costs
    milliseconds); an event guards the previous copy out of the buffer before it is overwritten."""

    def __init__(self):
        self.buf, self.evt = {}, {}

    def upload(self, key, arr):
        """Copy a contiguous numpy array to a new device tensor through the pinned buffer `key`."""
        torch = _torch()
        a = np.ascontiguousarray(arr).reshape(-1).view(np.uint8)
        n = max(a.nbytes, 8)
        buf = self.buf.get(key)
        if buf is None or buf.numel() < n:
            buf = self.buf[key] = torch.empty(max(n, 1 << 16) * 5 // 4, dtype=torch.uint8, pin_memory=True)
        elif key in self.evt:
            self.evt[key].synchronize()
        buf.numpy()[:a.nbytes] = a
        dev = torch.empty(n, dtype=torch.uint8, device="cuda")
        dev[:a.nbytes].copy_(buf[:a.nbytes], non_blocking=True)
        ev = self.evt[key] = torch.cuda.Event()
        ev.record()
        TRANSFER["h2d"] += a.nbytes
        return dev


_STAGE = _Staging()


def _scalar_min(name, v):
    if v is None:
        return float("nan")
    if np.ndim(v) != 0:
        raise NotImplementedError(f"find_peaks: `{name}` must be a scalar minimum (got an interval or an array)")
    return float(v)


class PeakBatch:
    """A ragged batch of spectra on the device: ``find`` then ``refine`` (or ``refine`` from
    caller-supplied peaks).  One host->device copy of the spectra at construction, one
    device->host copy per ``fetch_*``."""

    def __init__(self, spectra):
        self.lib = N.load()
        torch = _torch()
        self.spectra = [np.ascontiguousarray(s, dtype=np.float64).reshape(-1) for s in spectra]
        self.n = len(self.spectra)
        if self.n == 0:
            raise ValueError("empty batch")
        self.n_pix = np.array([len(s) for s in self.spectra], dtype=np.int64)
        self.pix_off = np.concatenate(([0], np.cumsum(self.n_pix)))
        total = int(self.pix_off[-1])
        flat = np.concatenate(self.spectra) if total else np.zeros(1)
        self.d_spec = _STAGE.upload("spectra", flat).view(torch.float64)
        self._found = None
        self._refined = None

    # ------------------------------------------------------------------ find_peaks
    def find(self, height=None, distance=None, prominence=None, max_peaks=None):
        """scipy.signal.find_peaks(s, height=, distance=, prominence=) for every spectrum.
        max_peaks: per-spectrum output capacity (default: the most a spectrum can hold)."""
        torch = _torch()
        h, dd, p = _scalar_min("height", height), _scalar_min("distance", distance), _scalar_min("prominence", prominence)
        if distance is not None and not dd >= 1:
            raise ValueError("`distance` must be greater or equal to 1")
        cap = self.n_pix // 2 + 1
        if max_peaks is not None:
            cap = np.minimum(cap, int(max_peaks))
        off = np.concatenate(([0], np.cumsum(cap)))
        tot = max(int(off[-1]), 1)
        f = dict(cap=cap, off=off, has_prom=prominence is not None,
                 peaks=torch.empty(tot, dtype=torch.int32, device="cuda"),
                 peaks_f64=torch.empty(tot, dtype=torch.float64, device="cuda"),
                 heights=torch.empty(tot, dtype=torch.float64, device="cuda"),
                 prom=torch.empty(tot if prominence is not None else 1, dtype=torch.float64, device="cuda"),
                 count=torch.zeros(self.n, dtype=torch.int32, device="cuda"))
        desc = np.zeros(self.n, dtype=np.dtype(N.FindPeaksDesc))
        desc["d_spectrum"] = self.d_spec.data_ptr() + 8 * self.pix_off[:-1]
        desc["n_pix"], desc["max_peaks"] = self.n_pix, cap
        desc["height"], desc["distance"], desc["prominence"] = h, dd, p
        desc["d_peaks"] = f["peaks"].data_ptr() + 4 * off[:-1]
        desc["d_peaks_f64"] = f["peaks_f64"].data_ptr() + 8 * off[:-1]
        desc["d_heights"] = f["heights"].data_ptr() + 8 * off[:-1]
        desc["d_prominences"] = (f["prom"].data_ptr() + 8 * off[:-1]) if prominence is not None else 0
        desc["d_n_peaks"] = f["count"].data_ptr() + 4 * np.arange(self.n)
        d_desc = self._upload(desc)
        rc = self.lib.rb2_find_peaks(self.n, d_desc.data_ptr(), desc.ctypes.data, current_stream_ptr())
        N.check(rc, "rb2_find_peaks")
        f["desc"] = (desc, d_desc)
        self._found = f
        return self

    def fetch_found(self):
        """[(peaks int64, properties dict)] per spectrum, as scipy.signal.find_peaks returns them."""
        f = self._found
        cnt = f["count"].cpu().numpy()
        if np.any(cnt > f["cap"]):
            i = int(np.argmax(cnt > f["cap"]))
            raise N.NativeError(f"find_peaks: spectrum {i} has {cnt[i]} peaks, capacity max_peaks={int(f['cap'][i])}")
        pk, ht = f["peaks"].cpu().numpy(), f["heights"].cpu().numpy()
        pr = f["prom"].cpu().numpy() if f["has_prom"] else None
        TRANSFER["d2h"] += cnt.nbytes + pk.nbytes + ht.nbytes + (pr.nbytes if pr is not None else 0)
        out = []
        for i in range(self.n):
            o, c = int(f["off"][i]), int(cnt[i])
            props = {"peak_heights": ht[o:o + c].copy()}
            if pr is not None:
                props["prominences"] = pr[o:o + c].copy()
            out.append((pk[o:o + c].astype(np.int64), props))
        return out

    # ---------------------------------------------------------------- refine_peaks
    def refine(self, window_width=10, peaks=None):
        """rascal.util.refine_peaks for every spectrum.  peaks=None: refine what ``find`` left on
        the device; otherwise a list of per-spectrum initial estimates (any real numbers)."""
        torch = _torch()
        ww = int(window_width)
        if not 1 <= ww <= N.REFINE_MAX_WINDOW:
            raise ValueError(f"window_width must be in 1..{N.REFINE_MAX_WINDOW}")
        if peaks is None:
            if self._found is None:
                raise ValueError("refine(peaks=None) needs a previous find()")
            f = self._found
            cap, off = f["cap"], f["off"]
            d_peaks, d_cnt = f["peaks_f64"], f["count"]
        else:
            if len(peaks) != self.n:
                raise ValueError("one list of peaks per spectrum")
            pl = [np.asarray(p, dtype=np.float64).reshape(-1) for p in peaks]
            cap = np.array([len(p) for p in pl], dtype=np.int64)
            off = np.concatenate(([0], np.cumsum(cap)))
            flat = np.concatenate(pl) if off[-1] else np.zeros(1)
            d_peaks, d_cnt = _STAGE.upload("peaks", flat).view(torch.float64), None
        tot = max(int(off[-1]), 1)
        r = dict(cap=cap, off=off, ww=ww, d_peaks=d_peaks,
                 work=torch.empty(max(int(self.pix_off[-1]), 1), dtype=torch.float64, device="cuda"),
                 window=torch.empty(tot * 2 * ww, dtype=torch.float64, device="cuda"),
                 fit=torch.empty(tot * 4, dtype=torch.float64, device="cuda"),
                 status=torch.zeros(tot, dtype=torch.int32, device="cuda"),
                 refined=torch.empty(tot, dtype=torch.float64, device="cuda"),
                 count=torch.zeros(self.n, dtype=torch.int32, device="cuda"))
        desc = np.zeros(self.n, dtype=np.dtype(N.RefineDesc))
        desc["d_spectrum"] = self.d_spec.data_ptr() + 8 * self.pix_off[:-1]
        desc["d_work"] = r["work"].data_ptr() + 8 * self.pix_off[:-1]
        desc["n_pix"], desc["n_peaks"], desc["window_width"] = self.n_pix, cap, ww
        desc["d_peaks"] = d_peaks.data_ptr() + 8 * off[:-1]
        desc["d_n_peaks"] = (d_cnt.data_ptr() + 4 * np.arange(self.n)) if d_cnt is not None else 0
        desc["d_window"] = r["window"].data_ptr() + 8 * 2 * ww * off[:-1]
        desc["d_fit"] = r["fit"].data_ptr() + 8 * 4 * off[:-1]
        desc["d_status"] = r["status"].data_ptr() + 4 * off[:-1]
        desc["d_refined"] = r["refined"].data_ptr() + 8 * off[:-1]
        desc["d_n_refined"] = r["count"].data_ptr() + 4 * np.arange(self.n)
        d_desc = self._upload(desc)
        r["queue"] = torch.empty(int(self.lib.rb2_refine_scratch_bytes(self.n)), dtype=torch.uint8, device="cuda")
        rc = self.lib.rb2_refine_peaks(self.n, d_desc.data_ptr(), desc.ctypes.data, r["queue"].data_ptr(),
                                       current_stream_ptr())
        N.check(rc, "rb2_refine_peaks")
        r["desc"] = (desc, d_desc)
        r["d_cnt_in"] = d_cnt
        self._refined = r
        return self

    def fetch_refined(self, details=False):
        """[refined peaks] per spectrum.  Raises what the reference raises for a window it cannot
        fit (empty: ValueError, fewer than three samples: TypeError).  details=True also returns
        per-spectrum dicts with the per-peak status names, lmdif ier / nfev and fitted parameters."""
        r = self._refined
        cnt = r["count"].cpu().numpy()
        ref = r["refined"].cpu().numpy()
        st = r["status"].cpu().numpy()
        TRANSFER["d2h"] += cnt.nbytes + ref.nbytes + st.nbytes
        n_in = r["cap"]
        if r["d_cnt_in"] is not None:
            n_in = r["d_cnt_in"].cpu().numpy()
            if np.any(n_in > r["cap"]):
                raise N.NativeError("find_peaks found more peaks than max_peaks holds; raise max_peaks")
        fit = r["fit"].cpu().numpy().reshape(-1, 4) if details else None
        out, det = [], []
        for i in range(self.n):
            o, k = int(r["off"][i]), int(n_in[i])
            code = st[o:o + k] & 0xFF
            if np.any(code == 6):
                raise ValueError("zero-size array to reduction operation fmax which has no identity "
                                 f"(spectrum {i}: a peak whose window is empty)")
            if np.any(code == 7):
                raise TypeError("The number of func parameters=3 must not exceed the number of data points "
                                f"(spectrum {i}: a window with fewer than three samples)")
            out.append(ref[o:o + int(cnt[i])].copy())
            if details:
                det.append(dict(status=[N.PEAK_STATUS[c] for c in code], ier=(st[o:o + k] >> 8),
                                nfev=fit[o:o + k, 3].astype(np.int64), height=fit[o:o + k, 0].copy(),
                                centre=fit[o:o + k, 1].copy(), sigma=fit[o:o + k, 2].copy()))
        return (out, det) if details else out

    def _upload(self, desc):
        return _STAGE.upload("desc", desc.view(np.uint8))


def find_peaks(x, height=None, threshold=None, distance=None, prominence=None, **kwargs):
    """``scipy.signal.find_peaks`` (height / distance / prominence) on the device:
    returns ``(peaks, properties)`` with ``peak_heights`` and, if asked, ``prominences``
    (SciPy's ``left_bases`` / ``right_bases`` are not reported; the reference's recipes
    discard the properties)."""
    if threshold is not None or any(kwargs.get(k) is not None for k in _UNSUPPORTED):
        raise NotImplementedError("find_peaks: only height, distance and prominence are implemented on the device")
    bad = set(kwargs) - set(_UNSUPPORTED)
    if bad:
        raise TypeError(f"find_peaks() got an unexpected keyword argument '{sorted(bad)[0]}'")
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError("`x` must be a 1-D array")
    peaks, props = PeakBatch([x]).find(height, distance, prominence).fetch_found()[0]
    if height is None:
        props.pop("peak_heights")
    return peaks, props


def refine_peaks(spectrum, peaks, window_width=10, distance=None):
    """Drop-in for ``rascal.util.refine_peaks`` (util.py:450-523); ``distance`` is accepted and
    ignored exactly as the reference does."""
    return PeakBatch([spectrum]).refine(window_width, [peaks]).fetch_refined()[0]


def find_and_refine(spectra, height=None, distance=None, prominence=None, window_width=10, max_peaks=None):
    """find_peaks + refine_peaks for a batch of spectra; the found peaks never leave the device.
    Returns the list of refined peak arrays (what ``Calibrator`` / ``calibrate_batch`` take)."""
    return PeakBatch(spectra).find(height, distance, prominence, max_peaks).refine(window_width).fetch_refined()
