"""`rascal.calibrator.Calibrator` with the calibration hot path on the B200.

Mirror of the reference class (src/rascal/calibrator.py:20-2324) for the path
this repo accelerates: same constructor, setters, defaults, attributes, return
tuples and error behaviour, so user code that does

    c = Calibrator(peaks, spectrum)
    c.set_hough_properties(...); c.set_ransac_properties(...)
    c.set_atlas(atlas); c.do_hough_transform(); c.fit(max_tries=..., fit_deg=...)

runs unchanged.  Host Python keeps what the reference keeps on the host
(property setters :859-1283, pair enumeration :86-133, the RANSAC set-up
:441-523, result assembly :1673-1715); the arithmetic runs in hand-written
sm_100a kernels behind the C ABI (include/rascal_b200.h):

    do_hough_transform  -> K1 rb2_hough_accumulate   (houghtransform.py:60-92, 167-210)
    fit                 -> K2 rb2_candidate_vote     (calibrator.py:154-261)
                           K3 rb2_ransac_batch       (calibrator.py:525-735)
                           K4 rb2_refit_winners      (calibrator.py:636-720, models.py:70-123)

There is no CPU fallback; without the built library or a CUDA device every hot
call raises.
"""
from __future__ import annotations

import copy
import logging

import numpy as np
from scipy import interpolate

from . import engine
from .houghtransform import HoughTransform

__all__ = ["Calibrator", "HoughTransform", "LineList"]


class LineList:
    """Minimal stand-in for rascal.atlas.Atlas holding user-supplied lines
    (atlas.py:206-281 without the air/vacuum conversion).  Any object with
    ``get_lines()`` and ``__len__`` - including the reference's own Atlas -
    can be passed to ``Calibrator.set_atlas``."""

    def __init__(self, wavelengths, elements=None, intensities=None):
        self._lines = [float(w) for w in wavelengths]
        self._elements = list(elements) if elements is not None else ["?"] * len(self._lines)
        self._intensities = list(intensities) if intensities is not None else [0] * len(self._lines)

    def get_lines(self):
        return list(self._lines)

    def get_elements(self):
        return list(self._elements)

    def get_intensities(self):
        return list(self._intensities)

    def __len__(self):
        return len(self._lines)


def _keep(new, old, default):
    """The reference's setter rule: explicit value > previous value > default."""
    if new is not None:
        return new
    return default if old is None else old


class Calibrator:
    def __init__(self, peaks, spectrum=None):
        self.logger = None
        self.log_level = None
        self.peaks = copy.deepcopy(peaks)
        self.spectrum = copy.deepcopy(spectrum)
        self.atlas = None
        self.pix_known = None
        self.wave_known = None
        self._hough_done = False
        self._hough_lines_override = None
        self.hough_points = None
        self.ht = HoughTransform()
        self.pairs = []
        self.num_pix = None
        self.pixel_list = None
        self.plotting_library = None
        self.constrain_poly = None
        self.num_slopes = self.xbins = self.ybins = None
        self.min_wavelength = self.max_wavelength = None
        self.range_tolerance = self.linearity_tolerance = None
        self.sample_size = self.top_n_candidate = self.linear = self.filter_close = None
        self.ransac_tolerance = self.candidate_weighted = self.hough_weight = None
        self.minimum_matches = self.minimum_peak_utilisation = self.minimum_fit_error = None
        self.matched_peaks = []
        self.matched_atlas = []
        self.fit_coeff = None
        self.candidate_peak = self.candidate_arc = None
        self.ransac_counters = None
        self.seed = None
        self.dist_group = False
        self.set_calibrator_properties()
        self.set_hough_properties()
        self.set_ransac_properties()

    # ------------------------------------------------------------------ setters
    def set_calibrator_properties(self, num_pix=None, pixel_list=None, plotting_library=None, seed=None,
                                  logger_name="Calibrator", log_level="warning"):
        """calibrator.py:859-982 (plotting-library selection is not part of this package)."""
        self.logger = logging.getLogger(logger_name)
        self.logger.propagate = False
        level = logging.getLevelName(log_level.upper())
        self.logger.setLevel(level)
        self.log_level = level
        if len(self.logger.handlers) == 0:
            handler = logging.StreamHandler()
            handler.setFormatter(logging.Formatter(
                "[%(asctime)s] %(levelname)s [%(filename)s:%(lineno)d] %(message)s",
                datefmt="%a, %d %b %Y %H:%M:%S"))
            self.logger.addHandler(handler)
        if num_pix is not None:
            self.num_pix = num_pix
        elif self.num_pix is None:
            try:
                self.num_pix = len(self.spectrum)
            except Exception as e:
                self.logger.warning(e)
                self.logger.warning("Neither num_pix nor spectrum is given, it uses 1.1 times max(peaks) as "
                                    "the maximum pixel value.")
                self.num_pix = 1.1 * max(self.peaks)
        if pixel_list is not None:
            self.pixel_list = np.asarray(pixel_list)
        elif self.pixel_list is None:
            self.pixel_list = np.arange(self.num_pix)
        self.__dict__["_pix_to_rawpix"] = None  # (:941-946) built when somebody asks: see the property
        if seed is not None:
            self.seed = seed
            np.random.seed(seed)
        self.plotting_library = _keep(plotting_library, self.plotting_library, "matplotlib")

    @property
    def pix_to_rawpix(self):
        """interp1d(pixel_list -> raw pixel index) (calibrator.py:941-946), built on first use: only the
        reference's plotting and effective-pixel helpers read it, a fit never does."""
        f = self.__dict__.get("_pix_to_rawpix")
        if f is None:
            f = self.__dict__["_pix_to_rawpix"] = interpolate.interp1d(self.pixel_list, np.arange(len(self.pixel_list)),
                                                                       fill_value="extrapolate")
        return f

    def set_hough_properties(self, num_slopes=None, xbins=None, ybins=None, min_wavelength=None,
                             max_wavelength=None, range_tolerance=None, linearity_tolerance=None):
        """calibrator.py:984-1119."""
        self.num_slopes = int(_keep(num_slopes, self.num_slopes, 2000))
        self.xbins = _keep(xbins, self.xbins, 100)
        self.ybins = _keep(ybins, self.ybins, 100)
        self.min_wavelength = _keep(min_wavelength, self.min_wavelength, 3000.0)
        self.max_wavelength = _keep(max_wavelength, self.max_wavelength, 9000.0)
        self.range_tolerance = _keep(range_tolerance, self.range_tolerance, 500)
        self.linearity_tolerance = _keep(linearity_tolerance, self.linearity_tolerance, 100)
        # start wavelength +/- tolerance; slope limits over the detector span (:1088-1116)
        self.min_intercept = self.min_wavelength - self.range_tolerance
        self.max_intercept = self.min_wavelength + self.range_tolerance
        span = np.ptp(self.pixel_list)
        self.min_slope = ((self.max_wavelength - self.range_tolerance - self.linearity_tolerance)
                          - (self.min_intercept + self.range_tolerance + self.linearity_tolerance)) / span
        self.max_slope = ((self.max_wavelength + self.range_tolerance + self.linearity_tolerance)
                          - (self.min_intercept - self.range_tolerance - self.linearity_tolerance)) / span
        if self.atlas is not None:
            self._generate_pairs()

    def set_ransac_properties(self, sample_size=None, top_n_candidate=None, linear=None, filter_close=None,
                              ransac_tolerance=None, candidate_weighted=None, hough_weight=None,
                              minimum_matches=None, minimum_peak_utilisation=None, minimum_fit_error=None):
        """calibrator.py:1121-1283."""
        self.sample_size = _keep(sample_size, self.sample_size, 5)
        self.top_n_candidate = _keep(top_n_candidate, self.top_n_candidate, 5)
        self.linear = _keep(linear, self.linear, True)
        self.filter_close = _keep(filter_close, self.filter_close, False)
        self.ransac_tolerance = _keep(ransac_tolerance, self.ransac_tolerance, 5)
        self.candidate_weighted = _keep(candidate_weighted, self.candidate_weighted, True)
        self.hough_weight = _keep(hough_weight, self.hough_weight, 1.0)
        if minimum_matches is not None:
            assert minimum_matches > 0
        self.minimum_matches = _keep(minimum_matches, self.minimum_matches, 0)
        if minimum_peak_utilisation is not None:
            assert minimum_peak_utilisation >= 0 and minimum_peak_utilisation <= 1.0
        self.minimum_peak_utilisation = _keep(minimum_peak_utilisation, self.minimum_peak_utilisation, 0)
        if minimum_fit_error is not None:
            assert minimum_fit_error >= 0
        self.minimum_fit_error = _keep(minimum_fit_error, self.minimum_fit_error, 1e-4)

    def use_distributed(self, group=None):
        """Shard the RANSAC hypotheses of fit() over the ranks of a torch.distributed process
        group (None = the default group); one process per GPU."""
        self.dist_group = group

    # ------------------------------------------------------------------ atlas / pairs
    def _generate_pairs(self):
        """calibrator.py:86-133: peak-major Cartesian product, optional polygon mask."""
        peaks = np.asarray(self.peaks, dtype=np.float64)
        lines = np.asarray(self.atlas.get_lines(), dtype=np.float64)
        pairs = np.column_stack((np.repeat(peaks, len(lines)), np.tile(lines, len(peaks))))  # itertools.product order
        if self.constrain_poly:
            from scipy.spatial import Delaunay

            px = self.pixel_list.max()
            valid_area = Delaunay([
                (0, self.max_intercept + self.candidate_tolerance),
                (0, self.min_intercept - self.candidate_tolerance),
                (px, self.max_wavelength - self.range_tolerance - self.candidate_tolerance),
                (px, self.max_wavelength + self.range_tolerance + self.candidate_tolerance)])
            mask = valid_area.find_simplex(pairs) >= 0
            self.pairs = pairs[mask]
        else:
            self.pairs = pairs

    def set_atlas(self, atlas, candidate_tolerance=10.0, constrain_poly=False):
        """calibrator.py:1399-1423."""
        self.atlas = atlas
        self.candidate_tolerance = candidate_tolerance
        self.constrain_poly = constrain_poly
        self._generate_pairs()

    def set_known_pairs(self, pix=(), wave=()):
        """calibrator.py:1507-1550."""
        pix = np.asarray(pix, dtype="float").reshape(-1)
        wave = np.asarray(wave, dtype="float").reshape(-1)
        assert pix.size == wave.size, ValueError(
            "Please check the length of the input arrays. pix has size {} and wave has size {}.".format(
                pix.size, wave.size))
        if not all(isinstance(p, (float, int)) & (not np.isnan(p)) for p in pix):
            raise ValueError("All pix elements have to be numeric.")
        if not all(isinstance(w, (float, int)) & (not np.isnan(w)) for w in wave):
            raise ValueError("All wave elements have to be numeric.")
        self.pix_known = pix
        self.wave_known = wave

    # ------------------------------------------------------------------ Hough
    def do_hough_transform(self, brute_force=False):
        """calibrator.py:1425-1452."""
        if len(self.pairs) == 0:
            logging.warning("pairs list is empty. Try generating now.")
            self._generate_pairs()
            if len(self.pairs) == 0:
                logging.error("pairs list is still empty.")
        self.ht.set_constraints(self.min_slope, self.max_slope, self.min_intercept, self.max_intercept)
        if brute_force:
            self.ht.generate_hough_points_brute_force(self.pairs[:, 0], self.pairs[:, 1])
        else:
            self.ht.generate_hough_points(self.pairs[:, 0], self.pairs[:, 1], num_slopes=self.num_slopes)
        self.ht.bin_hough_points(self.xbins, self.ybins)  # runs on the device when its results are first read
        self.hough_points = _LazyPoints(self.ht)
        self._hough_lines_override = None
        self._hough_done = True

    def save_hough_transform(self, filename="hough_transform", fileformat="npy", delimiter="+", to_disk=True):
        """calibrator.py:1454-1481."""
        self.ht.save(filename=filename, fileformat=fileformat, delimiter=delimiter, to_disk=to_disk)

    def load_hough_transform(self, filename="hough_transform", filetype="npy"):
        """calibrator.py:1483-1505: the loaded points are re-binned by the next do_hough_transform-less fit
        through HoughTransform.bin_hough_points (explicit-points path)."""
        self.ht.load(filename=filename, filetype=filetype)

    @property
    def hough_lines(self):
        """Bin-centre lines in rank order (calibrator.py:1450-1452), from the HoughTransform on first use."""
        if self._hough_lines_override is not None:
            return self._hough_lines_override
        return self.ht.hough_lines if self._hough_done else None

    @hough_lines.setter
    def hough_lines(self, value):
        self._hough_lines_override = value

    # ------------------------------------------------------------------ fit
    def fit(self, max_tries=500, fit_deg=4, fit_coeff=None, fit_tolerance=5.0, fit_type="poly",
            candidate_tolerance=2.0, brute_force=False, progress=True, n_hypotheses=None):
        """calibrator.py:1552-1715; returns (fit_coeff, matched_peaks, matched_atlas, rms, residual,
        peak_utilisation, atlas_utilisation).

        Extension: `n_hypotheses` (keyword) replaces the `max_tries` stopping rule by an exact
        budget of drawn hypotheses (ids 0..n-1 of the Philox stream), sharded over the ranks of
        `use_distributed()`."""
        self.max_tries = max_tries
        self.fit_deg = fit_deg
        self.fit_coeff = fit_coeff
        if fit_coeff is not None:
            self.fit_deg = len(fit_coeff) - 1
        self.fit_tolerance = fit_tolerance
        self.fit_type = fit_type
        self.brute_force = brute_force
        self.progress = progress
        if self.fit_type == "poly":
            self.polyfit = np.polynomial.polynomial.polyfit
            self.polyval = np.polynomial.polynomial.polyval
        elif self.fit_type == "legendre":
            self.polyfit = np.polynomial.legendre.legfit
            self.polyval = np.polynomial.legendre.legval
        elif self.fit_type == "chebyshev":
            self.polyfit = np.polynomial.chebyshev.chebfit
            self.polyval = np.polynomial.chebyshev.chebval
        else:
            raise ValueError("fit_type must be: (1) poly, (2) legendre or (3) chebyshev")
        if brute_force:
            raise NotImplementedError("brute_force sampling is not supported (it never terminates on a "
                                      "rejected sample in the reference, calibrator.py:532-534)")
        if self.sample_size > len(self.atlas):
            self.logger.warning("Size of sample_size is larger than the size of atlas, the sample_size is set to "
                                "match the size of atlas = " + str(len(self.atlas)) + ".")
            self.sample_size = len(self.atlas)
        # the reference compares with the fit_deg ARGUMENT (:1651) even when fit_coeff has just set self.fit_deg; where
        # that would leave a sample smaller than the fitted degree + 1 (a rank-deficient polyfit there) the degree
        # actually fitted decides
        need = max(int(fit_deg), int(self.fit_deg))
        if self.sample_size <= need:
            self.sample_size = need + 1
        if not self._hough_done:
            self.do_hough_transform()
        if self.minimum_matches > len(self.atlas):
            self.logger.warning("Requested minimum matches is greater than the atlas size, setting the minimum "
                                "number of matches to equal the atlas size = " + str(len(self.atlas)) + ".")
            self.minimum_matches = len(self.atlas)
        if self.minimum_matches > len(self.peaks):
            self.logger.warning("Requested minimum matches is greater than the number of peaks detected, which "
                                "has a size of size = " + str(len(self.peaks)) + ".")
            self.minimum_matches = len(self.peaks)

        fit_coeff, rms, residual, n_inliers, valid = self._solve_candidate_ransac(
            fit_deg=self.fit_deg, fit_coeff=self.fit_coeff, max_tries=self.max_tries,
            candidate_tolerance=candidate_tolerance, n_hypotheses=n_hypotheses)

        peak_utilisation = n_inliers / len(self.peaks)
        atlas_utilisation = n_inliers / len(self.atlas)
        if not valid:
            self.logger.warning("Invalid fit")
        if rms > self.fit_tolerance:
            self.logger.warning("RMS too large {} > {}".format(rms, self.fit_tolerance))
        assert fit_coeff is not None, "Couldn't fit"
        self.fit_coeff = fit_coeff
        self.rms = rms
        self.residual = residual
        self.peak_utilisation = peak_utilisation
        self.atlas_utilisation = atlas_utilisation
        return (self.fit_coeff, self.matched_peaks, self.matched_atlas, self.rms, self.residual,
                self.peak_utilisation, self.atlas_utilisation)

    def match_peaks(self, fit_coeff=None, n_delta=None, refine=False, tolerance=10.0, method="Nelder-Mead",
                    convergence=1e-6, min_frac=0.5, robust_refit=True, fit_deg=None, refit_mode=0):
        """calibrator.py:1717-1956 on the device (K5, rb2_match_peaks): every peak is associated
        with the atlas lines within `tolerance` of the given solution, the association is
        polyfit-ted and (robust_refit) re-fitted with models.robust_polyfit.

        Where the reference is defined (no peak has two atlas lines inside the tolerance) the
        7-tuple is the reference's.  Where it raises (an ambiguous peak: TypeError at :1867, see
        DESIGN.md) the nearest line is taken and a warning is logged.  `refine=True`: the reference's
        experimental Nelder-Mead pre-step (:1797-1821) over its objective _adjust_polyfit (:754-820), which runs
        on the device (rb2_adjust_polyfit); as in the reference its result only reaches the output when
        robust_refit=False - the association itself is made with the unrefined `fit_coeff` (:1831)."""
        if fit_coeff is None:
            fit_coeff = self.fit_coeff.copy()
        if fit_deg is None:
            fit_deg = len(fit_coeff) - 1
        if getattr(self, "polyval", None) is None:
            self.polyfit = np.polynomial.polynomial.polyfit
            self.polyval = np.polynomial.polynomial.polyval
        fit_coeff_new = None
        if refine:
            # :1797-1821 (experimental in the reference): Nelder-Mead on the host, exactly the reference's
            # scipy.optimize.minimize call, over Calibrator._adjust_polyfit evaluated on the device
            from scipy.optimize import minimize

            fit_coeff = np.asarray(fit_coeff, dtype=np.float64)
            fit_coeff_new = fit_coeff.copy()
            if n_delta is None:
                n_delta = len(fit_coeff_new) - 1
            objective = engine.AdjustPolyfit(self.peaks, self.atlas.get_lines(), self.pixel_list, fit_coeff, tolerance,
                                             min_frac, fit_type=getattr(self, "fit_type", None) or "poly")
            fitted_delta = minimize(objective, fit_coeff_new[: int(n_delta)] * 1e-3, method=method, tol=convergence,
                                    options={"maxiter": 10000}).x
            for i, d in enumerate(fitted_delta):
                fit_coeff_new[i] += d
            self.refine_evaluations = objective.n_evals
            if np.any(np.isnan(fit_coeff_new)):
                self.logger.warning("_adjust_polyfit() returns None. Input solution is returned.")
                return fit_coeff, None, None, None, None, None, None
        mb = engine.MatchBatch([dict(peaks=np.asarray(self.peaks, dtype=np.float64),
                                     lines=np.asarray(self.atlas.get_lines(), dtype=np.float64),
                                     coeffs=fit_coeff, tolerance=tolerance, fit_deg=fit_deg,
                                     robust_refit=robust_refit,
                                     fit_type=getattr(self, "fit_type", None) or "poly")]).run(mode=refit_mode)
        res, mx, my, rs = mb.fetch()
        r = res[0]
        m = r.n_matched
        if m == 0:
            raise TypeError("expected non-empty vector for x")  # np.polynomial.polyfit([], [], deg), :1893
        if m <= fit_deg:
            raise ValueError("match_peaks: %d matched peaks cannot constrain a degree-%d fit" % (m, fit_deg))
        if r.n_ambiguous:
            self.logger.warning("%d peaks have more than one atlas line within the tolerance; the nearest "
                                "line is used (the reference raises here)." % r.n_ambiguous)
        self.matched_peaks = mx[0, :m].copy()
        self.matched_atlas = my[0, :m].copy()
        self.match_ambiguous = int(r.n_ambiguous)
        self.peak_utilisation = m / len(self.peaks)
        self.atlas_utilisation = m / len(self.atlas)
        self.residuals = rs[0, :m].copy()
        if robust_refit:
            self.fit_coeff = np.array(r.refit.coeffs[:fit_deg + 1])
            self.rms = r.refit.rms
            self.refit_nfev = r.refit.huber_iters
        else:
            # :1944: fit_coeff_new, which only exists after refine=True (UnboundLocalError otherwise in the
            # reference; the documented intent is "the current solution": then the input is kept)
            self.fit_coeff = np.asarray(fit_coeff) if fit_coeff_new is None else fit_coeff_new
            self.rms = r.polyfit_rms
        return (self.fit_coeff, self.matched_peaks, self.matched_atlas, self.rms, self.residuals,
                self.peak_utilisation, self.atlas_utilisation)

    # candidate_peak / candidate_arc (calibrator.py:432-439): plain lists after the staged path; after the
    # device-resident path they are read back from K3's tables when somebody asks
    def _candidates(self, which):
        out = self.__dict__.get("_lazy_candidates")
        if out is not None:
            t = out.k3_tables(0)
            counts = np.diff(t["cand_off"])
            self.__dict__["_candidate_peak"] = list(np.repeat(t["peaks"], counts))
            self.__dict__["_candidate_arc"] = list(t["cand_wave"])
            self.__dict__["_lazy_candidates"] = None
        return self.__dict__.get(which)

    @property
    def candidate_peak(self):
        return self._candidates("_candidate_peak")

    @candidate_peak.setter
    def candidate_peak(self, v):
        self.__dict__["_lazy_candidates"] = None
        self.__dict__["_candidate_peak"] = v

    @property
    def candidate_arc(self):
        return self._candidates("_candidate_arc")

    @candidate_arc.setter
    def candidate_arc(self, v):
        self.__dict__["_lazy_candidates"] = None
        self.__dict__["_candidate_arc"] = v

    @property
    def candidates(self):
        """The reference's `self.candidates` (calibrator.py:239-261): one (pixels, wavelengths, weights) triple
        per Hough line in rank order - read by its plotting code (plotting.py:309-349).  The vote (K2) never
        builds this list; it is materialised from the device on first use after a fit."""
        c = self.__dict__.get("_cand_triples")
        if c is None and not self.linear and self.__dict__.get("_poly_candidates") is not None:
            # linear=False: one [peaks, lines, weights] entry per random shift that hit anything (:298-315)
            delta, base, actual, peaks, tol = self._poly_candidates
            c = []
            for d in delta:
                predicted = base + d
                mask = np.abs(actual - predicted) < tol
                if np.sum(mask) > 0:
                    w = np.exp(-((actual[mask] - predicted[mask]) ** 2) / (2 * self.range_tolerance ** 2 + 1e-9))
                    c.append([peaks[mask], actual[mask], w])
            self.__dict__["_cand_triples"] = c
        if c is None and self.__dict__.get("_cand_triples_tol") is not None:
            tol = self.__dict__["_cand_triples_tol"]
            sigma = (self.range_tolerance + self.linearity_tolerance) * 1.1775  # :256
            self.ht.hist  # noqa: B018  (runs the Hough transform if it is still pending)
            off, x, y, w = engine.candidate_points(self.ht._batch, 0, tol, 2 * sigma ** 2 + 1e-9, pairs=self.pairs)
            c = self.__dict__["_cand_triples"] = [(x[a:b], y[a:b], w[a:b]) for a, b in zip(off[:-1], off[1:])]
        return c

    @candidates.setter
    def candidates(self, v):
        self.__dict__["_cand_triples"] = v

    def _solve_device_resident(self, fit_deg, candidate_tolerance, n_hypotheses):
        """fit(n_hypotheses=N) without a host round trip between the stages: pairs -> K1 -> K2 ->
        rb2_ransac_prepare -> K3 (this rank's shard) -> all-gather + merge -> K4 in one chain
        (rascal_b200.pipeline), one H2D of peaks / lines / descriptors and one D2H of the result.
        Returns None when the problem does not qualify (known pairs, filter_close, polygon constraint,
        non-standard pixel list, ...) or the winner fails the refit-dependent acceptance - the staged
        path with its fall-through then takes over."""
        from . import parallel, pipeline

        if (self.pix_known is not None or self.constrain_poly or self.filter_close or not self.linear
                or self.ht._external or self.sample_size != fit_deg + 1 or getattr(self, "fit_type", "poly") != "poly"):
            return None
        pl = np.asarray(self.pixel_list, dtype=np.float64)
        # K1 / K2 / K3 work on the ascending distinct peaks whatever the caller's order (the staged path
        # reaches the same set through np.unique); duplicated peaks or unsorted atlas lines stay staged
        peaks = np.sort(np.asarray(self.peaks, dtype=np.float64))
        lines = np.asarray(self.atlas.get_lines(), dtype=np.float64)
        if not (len(pl) and pl[0] == 0 and pl[-1] == len(pl) - 1 and np.all(np.diff(pl) == 1)):
            return None
        spec = dict(peaks=peaks, lines=lines, pixel_list=pl, n_pix=len(pl), pix_is_arange=True, sample_size=self.sample_size,
                    limits=(self.min_slope, self.max_slope, self.min_intercept, self.max_intercept),
                    hough=dict(num_slopes=self.num_slopes, xbins=self.xbins, ybins=self.ybins,
                               range_tolerance=self.range_tolerance, linearity_tolerance=self.linearity_tolerance),
                    ransac=dict(top_n_candidate=self.top_n_candidate, candidate_weighted=self.candidate_weighted,
                                filter_close=False, ransac_tolerance=self.ransac_tolerance, hough_weight=self.hough_weight,
                                minimum_matches=self.minimum_matches,
                                minimum_peak_utilisation=self.minimum_peak_utilisation,
                                minimum_fit_error=self.minimum_fit_error),
                    fit=dict(fit_deg=fit_deg, candidate_tolerance=candidate_tolerance))
        if not pipeline.eligible([spec]):
            return None
        rank, world = parallel.rank_world(self.dist_group)
        key = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
        key = parallel.broadcast_int(key, self.dist_group)
        lo, hi = parallel.shard_range(n_hypotheses, rank, world)
        exchange = None
        if world > 1:
            def exchange(d_result, n):
                parallel.exchange_device(d_result, n, self.dist_group)
        out = pipeline.run([spec], hi - lo, seed=key, hyp_begin=lo, exchange=exchange, keep_tables=True)
        if out.law_error[0]:
            raise ValueError("probabilities contain NaN")  # np.random.choice in the reference's loop (:538)
        if not out.enough[0] or not out.ok[0]:
            return None  # nothing eligible, or the winner failed K4's acceptance: the staged path decides
        self._lazy_candidates = out  # candidate_peak / candidate_arc: unpacked on first use from tables `out` owns
        m = int(out.n_inliers[0])
        self.matched_peaks, self.matched_atlas = list(out.mx[0, :m]), list(out.my[0, :m])
        assert len(self.matched_atlas) == len(set(self.matched_atlas))
        self.ransac_counters = {k: int(out.counters[0, j]) for j, k in enumerate(engine.N.COUNTER_NAMES)}
        if self.ransac_counters["unsupported"]:
            # hypotheses outside the corner-clamped Hough-weight regime (nm-scale problems): the staged path
            # carries the spline tables of the general weight
            return None
        self.best_cost = float(out.cost[0])
        self.best_hypothesis = (int(out.hyp_id[0]), np.array(out.hyp_coeffs[0, :fit_deg + 1]))
        self.refit_nfev = int(out.refit_nfev[0])
        return (np.array(out.coeffs[0, :fit_deg + 1]), float(out.rms[0]), out.rs[0, :m].copy(), m, m > fit_deg + 1)

    def _vote(self, candidate_tolerance):
        """K2: _get_candidate_points_linear + _get_most_common_candidates (calibrator.py:154-261)."""
        sigma = (self.range_tolerance + self.linearity_tolerance) * 1.1775  # :256
        vb = engine.VoteBatch(self.ht._batch, [dict(
            pairs=self.pairs, top_n=self.top_n_candidate, weighted=self.candidate_weighted,
            tol=candidate_tolerance, gauss_denom=2 * sigma ** 2 + 1e-9)]).run()
        cp, ca = vb.candidates(0)
        self.candidate_peak, self.candidate_arc = list(cp), list(ca)

    def _span_tries(self, rb, setup, begin, count, need, best, key, rank, world):
        """Evaluate hypothesis ids [begin, begin + count) (this rank's shard of them) and count the tries of the
        reference loop they complete: -> (completed tries, length of the prefix that holds `need` of them - `count`
        if the span holds fewer -, merged winner record of the whole span).

        A try is completed by an impossible draw, by a scored hypothesis whose cost does not beat the best accepted
        so far, or by one that is accepted (here: passes the inlier-count criteria; the refit-dependent
        minimum_fit_error test is left to K4's fall-through).  The best-so-far at hypothesis i is the minimum of
        the accepted costs before it in id order: a prefix minimum on the device, seeded with `best` and, on rank
        r, with the minima of the ranks before it (one tiny all-gather)."""
        from . import parallel

        torch = engine._torch()
        N = engine.N
        lo, hi = parallel.shard_range(count, rank, world)
        out = rb.run(begin + lo, hi - lo, seed=key, per_hypothesis=("status", "cost", "counts"), device_outputs=True)
        rb.exchange(self.dist_group)
        span_win = rb.results()[0]
        st, cost, ni = out["status"], out["cost"], out["counts"][:, 1]
        # everything below works on the scored hypotheses only (~1 % of the draws): ids ascending, their cost,
        # whether they pass the inlier-count criteria
        sc_idx = torch.nonzero(st == N.HYP_SCORED).squeeze(1)
        sc_cost, sc_ni = cost[sc_idx], ni[sc_idx]
        sc_elig = (sc_ni >= setup.min_inliers) & (sc_ni.to(torch.float64) >= setup.min_inliers_frac) & (sc_cost <= setup.cost_ceiling)
        e_cost = sc_cost[sc_elig]
        start = best
        if world > 1:
            mins = parallel.allgather_doubles([float(e_cost.min().item()) if e_cost.numel() else float("inf")], self.dist_group)[:, 0]
            start = min([best] + [float(m) for m in mins[:rank]])
        # cm[k] = best accepted cost after the first k eligible hypotheses of this shard (a scan over ~1e5 values)
        cm = torch.cummin(torch.cat((torch.tensor([start], dtype=cost.dtype, device=cost.device), e_cost)), 0).values
        n_elig_before = torch.cumsum(sc_elig.to(torch.int64), 0) - sc_elig.to(torch.int64)
        beats = sc_cost <= cm[n_elig_before]
        done = ((st == N.HYP_ABORT) | (st == N.HYP_UNSUPPORTED)).to(torch.int64)
        done[sc_idx] = (~beats | sc_elig).to(torch.int64)
        mine = int(done.sum().item())
        counts = parallel.allgather_ints([mine], self.dist_group)[:, 0] if world > 1 else np.array([mine])
        total, before = int(counts.sum()), int(counts[:rank].sum())
        cut = count
        if total >= need and before < need <= before + mine:
            k = need - before  # the k-th completed try of this shard ends the budget
            pos = int(torch.searchsorted(torch.cumsum(done, 0), torch.tensor([k], device=done.device, dtype=torch.int64)).item())
            cut = lo + pos + 1
        if world > 1:
            cut = int(parallel.allgather_ints([cut], self.dist_group)[:, 0].min())
        return min(total, need), cut, span_win

    def _vote_poly(self, candidate_tolerance):
        """linear=False (experimental in the reference): _get_candidate_points_poly (calibrator.py:263-315) +
        _get_most_common_candidates (:154-217).  The reference compares peak i with atlas line i (its
        `actual - predicted` is elementwise: equal lengths or a broadcasting ValueError) for len(hough_lines)
        random shifts of the prior solution, so a peak's only possible candidate is its own line - kept iff
        some shift brings the prediction within the tolerance.  Set-up logic on the host like _generate_pairs;
        the shifts come from the seeded global NumPy RNG exactly as in the reference."""
        if self.fit_coeff is None:
            raise ValueError("A guess solution for a polynomial fit has to be provided as fit_coeff in fit() in "
                             "order to generate candidates for RANSAC sampling.")
        peaks = np.asarray(self.peaks, dtype=np.float64)
        actual = np.array(self.atlas.get_lines())
        n = int(self.xbins) * int(self.ybins)  # len(self.hough_lines)
        delta = np.random.random(n) * self.range_tolerance * 2.0 - self.range_tolerance
        base = self.polyval(peaks, self.fit_coeff)
        hit = np.abs(actual[None, :] - (base[None, :] + delta[:, None])) < candidate_tolerance  # [shift, peak]
        keep = hit.any(axis=0)
        order = np.argsort(peaks[keep], kind="stable")
        self.candidate_peak, self.candidate_arc = list(peaks[keep][order]), list(actual[keep][order])
        self._poly_candidates = (delta, base, actual, peaks, candidate_tolerance)

    def _solve_candidate_ransac(self, fit_deg, fit_coeff, max_tries, candidate_tolerance, n_hypotheses=None):
        """calibrator.py:382-752 as batched device work (order-independent rule).

        With torch.distributed initialised (one process per GPU) every rank evaluates its own
        slice of each hypothesis-id range and the per-rank winners are merged by ONE all-gather
        (rascal_b200.parallel); the result is identical on every rank and independent of the
        number of GPUs."""
        from . import parallel

        if not self.ht.binned:
            raise NotImplementedError("fit() needs a Hough transform computed by do_hough_transform()")
        self.__dict__["_cand_triples"], self.__dict__["_cand_triples_tol"] = None, candidate_tolerance
        if n_hypotheses is not None and fit_coeff is None:
            fast = self._solve_device_resident(fit_deg, candidate_tolerance, int(n_hypotheses))
            if fast is not None:
                return fast
        if self.linear:
            self._vote(candidate_tolerance)
        else:
            self._vote_poly(candidate_tolerance)
        best = (None, 1e50, None, 0, False)
        setup = engine.RansacSetup(
            self.candidate_peak, self.candidate_arc, fit_deg=fit_deg, sample_size=self.sample_size,
            ransac_tolerance=self.ransac_tolerance, min_intercept=self.min_intercept,
            max_intercept=self.max_intercept, pixel_list=self.pixel_list, hist=self.ht.hist,
            xedges=self.ht.xedges, yedges=self.ht.yedges, hough_weight=self.hough_weight,
            filter_close=self.filter_close, minimum_matches=self.minimum_matches,
            minimum_peak_utilisation=self.minimum_peak_utilisation, minimum_fit_error=self.minimum_fit_error,
            n_peaks_total=len(self.peaks), pix_known=self.pix_known, wave_known=self.wave_known,
            fit_type=getattr(self, "fit_type", "poly"))
        if not setup.enough:  # :468-469
            return best
        if fit_coeff is not None:  # :512-515: only hypotheses that beat the pre-existing fit count
            # One documented deviation: a prior that fits exact synthetic data perfectly has a summed error of
            # exactly 0.0, and the reference then only accepts a hypothesis whose LAPACK fit happens to hit the
            # same floating-point zero (its own test, test_synthetic_calibration.py:65-121, passes that way: 1 of
            # 100 tries).  Errors below 1e-9 A in total are rounding noise of whichever solver made them, so the
            # ceiling never drops below that.
            setup.cost_ceiling = max(setup.prior_cost(fit_coeff), 1e-9)
        if self.sample_size >= len(setup.peaks):
            raise NotImplementedError("sample_size >= number of candidate peaks switches the reference to its "
                                      "brute-force sampler (calibrator.py:474-477), which is not supported")
        rank, world = parallel.rank_world(self.dist_group)
        # Philox key: drawn from the legacy global RNG so that seed= makes the fit reproducible
        key = int(np.random.randint(0, 2 ** 31 - 1)) | (int(np.random.randint(0, 2 ** 31 - 1)) << 32)
        key = parallel.broadcast_int(key, self.dist_group)
        rb = engine.RansacBatch([setup])

        def evaluate(begin, count, floors=None):
            """K3 over this rank's slice + the one exchange; the merged winner stays on the device."""
            lo, hi = parallel.shard_range(count, rank, world)
            rb.run(begin + lo, hi - lo, seed=key, floors=floors)
            rb.exchange(self.dist_group)

        def one(struct):
            arr = engine.N.struct_array(engine.N.RansacResult, 1)
            arr[0] = struct
            return arr

        spans = []
        if n_hypotheses is not None:
            spans.append((0, int(n_hypotheses)))
            evaluate(0, int(n_hypotheses))
            rb.refit_async()  # K4 straight from the device-resident winner: one D2H for the whole fit
            win, fit, mx, my, rs = rb.fetch()
            winner = engine.N.RansacResult.from_buffer_copy(bytes(win[0]))
        else:
            # `max_tries` counts COMPLETED tries of the reference loop: an impossible draw, or a
            # scored hypothesis that either does not beat the best so far or is accepted; rejected draws and
            # hypotheses that beat the best but fail the acceptance checks are redrawn for free (:578-700).  Spans
            # of hypothesis ids are evaluated with their per-hypothesis outcomes kept on the device, where a prefix
            # minimum over the id order tells which ones completed a try; the last span is cut at the id where the
            # count reaches max_tries EXACTLY, so the outcome is a function of (seed, max_tries) only - not of the
            # span sizes or of the number of GPUs.
            winner = parallel.merge_winners([])
            done, chunk, completed = 0, max(1 << 14, 64 * int(max_tries)), 0
            while done < (1 << 40):
                best_so_far = min(setup.cost_ceiling, winner.cost if winner.hyp_id >= 0 else float("inf"))
                n_done, cut, span_win = self._span_tries(rb, setup, done, chunk, int(max_tries) - completed, best_so_far,
                                                         key, rank, world)
                if cut < chunk:  # the budget ends inside this span: its winner is the one of the prefix
                    evaluate(done, cut)
                    span_win = rb.results()[0]
                spans.append((done, cut))
                winner = parallel.merge_winners([winner, span_win])
                completed += n_done
                done += chunk
                if completed >= int(max_tries) or (span_win.counters[0] == 0 and span_win.counters[1] == 0):
                    break
                chunk = min(chunk * 4, 1 << 26)
            self.completed_tries = completed
            fit, mx, my, rs = rb.refit(one(winner))
        self.ransac_counters = engine.counters_dict(winner)
        if self.ransac_counters["unsupported"]:
            self.logger.warning("%d hypotheses left the corner-clamped Hough-weight regime and were skipped",
                                self.ransac_counters["unsupported"])
        # refit-dependent acceptance (:652-679); fall through to the next best on failure.  The reference keeps
        # drawing until a hypothesis passes; here every rejected winner costs one more pass over the spans
        # (noise-free data with a large minimum_fit_error rejects one perfect fit after the other)
        for _ in range(256):
            if winner.hyp_id < 0:
                return best
            o = fit[0]
            if o.accepted:
                m = o.n_inliers
                self.matched_peaks = list(mx[0, :m])
                self.matched_atlas = list(my[0, :m])
                assert len(self.matched_atlas) == len(set(self.matched_atlas))
                coeffs = np.array(o.coeffs[:fit_deg + 1])
                self.best_cost = winner.cost
                self.best_hypothesis = (winner.hyp_id, np.array(winner.coeffs[:fit_deg + 1]))
                self.refit_nfev = o.huber_iters
                return coeffs, o.rms, rs[0, :m].copy(), m, m > fit_deg + 1
            floor = (winner.cost, winner.hyp_id)
            nxt = parallel.merge_winners([])
            for b0, c0 in spans:
                evaluate(b0, c0, floors=[floor])
                nxt = parallel.merge_winners([nxt, rb.results()[0]])
            for k in range(8):
                nxt.counters[k] = winner.counters[k]
            winner = nxt
            fit, mx, my, rs = rb.refit(one(winner))
        return best


class _LazyPoints:
    """`Calibrator.hough_points`: materialised from the device on first use."""

    def __init__(self, ht):
        self._ht = ht

    def __array__(self, dtype=None, copy=None):
        a = self._ht.hough_points
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, k):
        return self._ht.hough_points[k]

    def __len__(self):
        return len(self._ht.hough_points)

    @property
    def shape(self):
        return self._ht.hough_points.shape
