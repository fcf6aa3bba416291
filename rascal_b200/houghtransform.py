"""`rascal.houghtransform.HoughTransform` with the hot methods on the B200.

Mirror of the reference class (src/rascal/houghtransform.py:5-339): same
method names, arguments, attributes and error behaviour.  The two hot methods,
`generate_hough_points` (:60-92) and `bin_hough_points` (:167-210), run as ONE
fused device pass (K1, rb2_hough_accumulate): the S*P intercept grid is never
materialised; `hough_points` is a lazy attribute produced by
rb2_hough_points in the reference's exact row order when somebody reads it.
"""
from __future__ import annotations

import json

import numpy as np

from . import engine


class HoughTransform:
    """Hough transform operations on the pixel-wavelength space."""

    def __init__(self):
        self._points = None          # materialised / externally supplied points
        self._pending = None         # (x, y, num_slopes) of a generate call not yet binned
        self._external = False       # points came from add_hough_points / load / brute force
        self._lazy_bins = None       # (xbins, ybins) of a bin_hough_points call not yet run on the device
        self._res = dict(hough_lines=None, hist=None, xedges=None, yedges=None, hist_sorted_arg=None, batch=None)
        self.min_slope = None
        self.max_slope = None
        self.min_intercept = None
        self.max_intercept = None

    # Results of bin_hough_points are produced on first use: a fit that runs device-resident end to end
    # (Calibrator.fit(n_hypotheses=...)) never needs the histogram on the host, and K1 then runs once,
    # inside that chain, instead of twice.
    def _get(self, name):
        if self._lazy_bins is not None:
            xb, yb = self._lazy_bins
            self._lazy_bins = None
            self._bin_now(xb, yb)
        if self._res[name] is None and name != "batch" and self._res["batch"] is not None:
            self._fetch(name)  # each host copy is made when somebody reads it (10^6 Hough lines are 16 MB of tuples)
        return self._res[name]

    def _fetch(self, name):
        batch = self._res["batch"]
        if name == "hist":
            self._res["hist"] = batch.hist(0).astype(np.float64)
        elif name == "xedges":
            self._res["xedges"] = batch.xedges(0)
        elif name == "yedges":
            self._res["yedges"] = batch.yedges(0)
        else:  # hist_sorted_arg / hough_lines: the ranking (houghtransform.py:190-210)
            xb, yb = batch._dims(0)
            i, j = np.unravel_index(batch.rank(0), (xb, yb))
            self._res["hist_sorted_arg"] = np.column_stack((i, j))
            xe, ye = self._get("xedges"), self._get("yedges")
            # bin-centre lines in rank order (:197-210); an [B,2] array instead of a list of tuples
            self._res["hough_lines"] = np.column_stack((xe[i] + (xe[1] - xe[0]) / 2, ye[j] + (ye[1] - ye[0]) / 2))

    def _set(self, name, value):
        self._res[name] = value

    hough_lines = property(lambda self: self._get("hough_lines"), lambda self, v: self._set("hough_lines", v))
    hist = property(lambda self: self._get("hist"), lambda self, v: self._set("hist", v))
    xedges = property(lambda self: self._get("xedges"), lambda self, v: self._set("xedges", v))
    yedges = property(lambda self: self._get("yedges"), lambda self, v: self._set("yedges", v))
    hist_sorted_arg = property(lambda self: self._get("hist_sorted_arg"), lambda self, v: self._set("hist_sorted_arg", v))
    _batch = property(lambda self: self._get("batch"), lambda self, v: self._set("batch", v))

    @property
    def binned(self):
        """True once bin_hough_points has been called (its device pass may still be pending)."""
        return self._lazy_bins is not None or self._res["batch"] is not None or self._res["hist"] is not None

    # ---- reference: houghtransform.py:22-58 --------------------------------
    def set_constraints(self, min_slope, max_slope, min_intercept, max_intercept):
        assert np.isfinite(min_slope), "min_slope has to be finite, %s is given " % min_slope
        assert np.isfinite(max_slope), "max_slope has to be finite, %s is given " % max_slope
        assert np.isfinite(min_intercept), "min_intercept has to be finite, %s is given " % min_intercept
        assert np.isfinite(max_intercept), "max_intercept has to be finite, %s is given " % max_intercept
        self.min_slope = min_slope
        self.max_slope = max_slope
        self.min_intercept = min_intercept
        self.max_intercept = max_intercept

    # ---- reference: houghtransform.py:60-92 --------------------------------
    def generate_hough_points(self, x, y, num_slopes):
        """Register the (pixel, wavelength) pairs and the slope grid.  The
        device pass runs in `bin_hough_points`; reading `hough_points` before
        that materialises them with a provisional 1x1 binning."""
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        self._pending = (x, y, int(num_slopes))
        self._points = None
        self._lazy_bins = None
        self._batch = None
        self._external = False

    # ---- reference: houghtransform.py:94-140 -------------------------------
    def generate_hough_points_brute_force(self, x, y):
        """The line through every two (pixel, wavelength) pairs, constrained strictly inside the
        slope / intercept limits, on the device (rb2_hough_brute_force)."""
        self.hough_points = engine.brute_force_points(x, y, self.min_slope, self.max_slope, self.min_intercept,
                                                      self.max_intercept)
        self._batch = None

    def _problem(self, xbins, ybins):
        x, y, S = self._pending
        return dict(pair_x=x, pair_y=y, min_slope=self.min_slope, max_slope=self.max_slope,
                    min_intercept=self.min_intercept, max_intercept=self.max_intercept,
                    n_slopes=S, xbins=int(xbins), ybins=int(ybins))

    @property
    def hough_points(self):
        if self._points is None and self._pending is not None:
            if self._batch is None:
                self._batch = engine.HoughBatch([self._problem(1, 1)]).run()
            self._points = self._batch.points(0)
        return self._points

    @hough_points.setter
    def hough_points(self, value):
        self._points = value
        self._external = value is not None
        if value is not None:
            self._pending = None

    # ---- reference: houghtransform.py:142-165 ------------------------------
    def add_hough_points(self, hp):
        if isinstance(hp, HoughTransform):
            points = hp.hough_points
        elif isinstance(hp, np.ndarray):
            points = hp
        else:
            raise TypeError("Unsupported type for extending hough points.")
        merged = np.vstack((self.hough_points, points))
        self._points = merged
        self._external = True
        self._pending = None

    # ---- reference: houghtransform.py:167-210 ------------------------------
    def bin_hough_points(self, xbins, ybins):
        assert (self._pending is not None) or (self._points is not None), \
            "Please load an hough_points or create an hough_points with hough_points() first."
        self._lazy_bins = (int(xbins), int(ybins))

    def _bin_now(self, xbins, ybins):
        if self._external or self._pending is None:  # brute force / add_hough_points / load: explicit points
            batch = engine.PointsHoughBatch([dict(points=self._points, xbins=int(xbins), ybins=int(ybins))]).run()
        else:
            batch = engine.HoughBatch([self._problem(xbins, ybins)]).run()
        for k in ("hist", "xedges", "yedges", "hist_sorted_arg", "hough_lines"):
            self._res[k] = None  # host copies are fetched from the batch on first use (_fetch)
        self._batch = batch

    # ---- reference: houghtransform.py:212-339 ------------------------------
    def save(self, filename="hough_transform", fileformat="npy", delimiter="+", to_disk=True):
        fmt = fileformat.split(delimiter)
        output_npy = output_json = None
        if "npy" in fmt:
            output_npy = [self.hough_points, self.hist, self.xedges, self.yedges, [self.min_slope],
                          [self.max_slope], [self.min_intercept], [self.max_intercept]]
            if to_disk:
                arr = np.empty(len(output_npy), dtype=object)
                for k, v in enumerate(output_npy):
                    arr[k] = v
                np.save(filename + ".npy", arr)
        if "json" in fmt:
            output_json = dict(hough_points=np.asarray(self.hough_points).tolist(), hist=self.hist.tolist(),
                               xedges=self.xedges.tolist(), yedges=self.yedges.tolist(),
                               min_slope=self.min_slope, max_slope=self.max_slope,
                               min_intercept=self.min_intercept, max_intercept=self.max_intercept)
            if to_disk:
                with open(filename + ".json", "w+") as f:
                    json.dump(output_json, f)
        if not to_disk:
            if output_npy is not None and output_json is not None:
                return output_npy, output_json
            return output_npy if output_npy is not None else output_json

    def load(self, filename="hough_transform", filetype="npy"):
        if filetype == "npy":
            if filename[-4:] != ".npy":
                filename += ".npy"
            z = np.load(filename, allow_pickle=True)
            pts, hist, xe, ye = z[0], z[1], z[2], z[3]
            lims = [float(z[k][0]) for k in range(4, 8)]
        elif filetype == "json":
            if filename[-5:] != ".json":
                filename += ".json"
            with open(filename) as f:
                z = json.load(f)
            pts, hist, xe, ye = z["hough_points"], z["hist"], z["xedges"], z["yedges"]
            lims = [float(z[k]) for k in ("min_slope", "max_slope", "min_intercept", "max_intercept")]
        else:
            raise ValueError("Unknown filetype %s, it has to be npy or json" % filetype)
        self.hough_points = np.asarray(pts, dtype=np.float64)
        self.hist = np.asarray(hist).astype("float")
        self.xedges = np.asarray(xe).astype("float")
        self.yedges = np.asarray(ye).astype("float")
        self.min_slope, self.max_slope, self.min_intercept, self.max_intercept = lims
