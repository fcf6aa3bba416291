"""ctypes binding of librascal_b200.so (the C ABI in include/rascal_b200.h).

There is no CPU fallback: importing a compute entry point without the built
library, or calling it without a CUDA device, raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RB2_LIB") or os.path.join(_HERE, "librascal_b200.so")  # RB2_LIB: a tuning variant (csrc/build.py)

ABI_VERSION = 2
RB2_MAX_DEG = 7
RB2_MAX_SAMPLE = 16
RB2_MAX_TOPN = 32
NCOEF = RB2_MAX_DEG + 1

_vp = C.c_void_p


class HoughDesc(C.Structure):
    _fields_ = [
        ("min_slope", C.c_double), ("max_slope", C.c_double),
        ("min_intercept", C.c_double), ("max_intercept", C.c_double),
        ("n_slopes", C.c_int32), ("xbins", C.c_int32), ("ybins", C.c_int32), ("n_pairs", C.c_int32),
        ("d_pair_x", _vp), ("d_pair_y", _vp), ("d_lo", _vp), ("d_hi", _vp),
        ("d_xedges", _vp), ("d_yedges", _vp), ("d_hist", _vp), ("d_rank", _vp),
        ("d_rankpos", _vp), ("d_sort_tmp", _vp),
    ]


class HoughStats(C.Structure):
    _fields_ = [
        ("icpt_min", C.c_double), ("icpt_max", C.c_double),
        ("grad_min", C.c_double), ("grad_max", C.c_double),
        ("n_points", C.c_int64), ("slope_lo", C.c_int32), ("slope_hi", C.c_int32),
        ("max_count", C.c_uint32), ("status", C.c_int32),
    ]


class VoteDesc(C.Structure):
    _fields_ = [
        ("n_upeaks", C.c_int32), ("n_uwaves", C.c_int32), ("xbins", C.c_int32), ("ybins", C.c_int32),
        ("top_n", C.c_int32), ("weighted", C.c_int32),
        ("tol", C.c_double), ("gauss_denom", C.c_double),
        ("d_upeak_x", _vp), ("d_pair_off", _vp), ("d_pair_y", _vp), ("d_pair_wid", _vp),
        ("d_xedges", _vp), ("d_yedges", _vp), ("d_rankpos", _vp),
        ("d_cand_wave", _vp), ("d_cand_pair", _vp), ("d_cand_n", _vp),
        ("d_agg", _vp), ("d_cnt", _vp), ("d_status", _vp), ("d_rank", _vp),
    ]


class RansacDesc(C.Structure):
    _fields_ = [
        ("n_upeaks", C.c_int32), ("n_wids", C.c_int32), ("n_cand", C.c_int32), ("pad0", C.c_int32),
        ("d_peaks", _vp), ("d_cand_off", _vp), ("d_cand_wave", _vp), ("d_cand_wid", _vp), ("d_cdf32", _vp),
        ("n_known", C.c_int32), ("d_known_x", _vp), ("d_known_y", _vp),
        ("deg", C.c_int32), ("sample_size", C.c_int32),
        ("min_intercept", C.c_double), ("max_intercept", C.c_double),
        ("pix_min", C.c_double), ("pix_max", C.c_double), ("tol", C.c_double),
        ("hough_weight", C.c_double), ("gc_min", C.c_double), ("gc_max", C.c_double),
        ("ic_min", C.c_double), ("h_lo", C.c_double), ("h_hi", C.c_double),
        ("n_pp", C.c_int32), ("d_pp_breaks", _vp), ("d_pp_coef", _vp),
        ("w_upper", C.c_double), ("n_pix", C.c_int32), ("d_pix", _vp),
        ("min_inliers", C.c_int32), ("min_inliers_frac", C.c_double),
        ("cost_ceiling", C.c_double), ("floor_cost", C.c_double), ("floor_id", C.c_int64),
        ("hyp_begin", C.c_int64), ("n_hyp", C.c_int64), ("seed", C.c_uint64),
        ("d_replay_x", _vp), ("d_replay_y", _vp),
        ("d_out_status", _vp), ("d_out_cost", _vp), ("d_out_weight", _vp),
        ("d_out_coeffs", _vp), ("d_out_counts", _vp), ("d_out_sample", _vp), ("d_replay_coeffs", _vp),
        ("min_fit_error", C.c_double),
        ("bs_nx", C.c_int32), ("bs_ny", C.c_int32), ("d_bs_tx", _vp), ("d_bs_ty", _vp), ("d_bs_c", _vp),
        ("d_bs_ppx", _vp), ("d_bs_ppy", _vp),
        ("basis", C.c_int32), ("pad1", C.c_int32), ("d_basis_c0", _vp), ("d_basis_grad", _vp),
    ]


class RansacResult(C.Structure):
    _fields_ = [
        ("cost", C.c_double), ("hyp_id", C.c_int64),
        ("n_match", C.c_int32), ("n_inliers", C.c_int32),
        ("coeffs", C.c_double * NCOEF), ("counters", C.c_uint64 * 8),
    ]


class RefitResult(C.Structure):
    _fields_ = [
        ("coeffs", C.c_double * NCOEF), ("rms", C.c_double),
        ("n_inliers", C.c_int32), ("accepted", C.c_int32),
        ("huber_iters", C.c_int32), ("pad", C.c_int32),
    ]


class MatchDesc(C.Structure):
    _fields_ = [
        ("d_peaks", _vp), ("d_atlas", _vp), ("n_peaks", C.c_int32), ("n_atlas", C.c_int32),
        ("coeffs", C.c_double * NCOEF), ("d_coeffs", _vp), ("n_coef", C.c_int32), ("fit_deg", C.c_int32),
        ("tolerance", C.c_double), ("robust_refit", C.c_int32), ("basis", C.c_int32),
    ]


class MatchResult(C.Structure):
    _fields_ = [
        ("n_matched", C.c_int32), ("n_ambiguous", C.c_int32),
        ("polyfit_coeffs", C.c_double * NCOEF), ("polyfit_err", C.c_double), ("polyfit_rms", C.c_double),
        ("refit", RefitResult),
    ]


class PairsDesc(C.Structure):
    _fields_ = [
        ("d_peaks", _vp), ("d_lines", _vp), ("n_peaks", C.c_int32), ("n_lines", C.c_int32),
        ("d_pair_x", _vp), ("d_pair_y", _vp), ("d_pair_wid", _vp), ("d_pair_off", _vp),
    ]


class PrepareDesc(C.Structure):
    _fields_ = [
        ("d_upeak_x", _vp), ("d_cand_wave", _vp), ("d_cand_n", _vp), ("n_peaks", C.c_int32), ("top_n", C.c_int32),
        ("d_xedges", _vp), ("d_yedges", _vp), ("d_hist", _vp), ("xbins", C.c_int32), ("ybins", C.c_int32),
        ("d_spline_op", _vp), ("d_spline_brk", _vp), ("n_pp", C.c_int32), ("pad", C.c_int32), ("d_enough", _vp),
    ]


class FindPeaksDesc(C.Structure):
    _fields_ = [
        ("d_spectrum", _vp), ("n_pix", C.c_int32), ("max_peaks", C.c_int32),
        ("height", C.c_double), ("distance", C.c_double), ("prominence", C.c_double),
        ("d_peaks", _vp), ("d_peaks_f64", _vp), ("d_heights", _vp), ("d_prominences", _vp), ("d_n_peaks", _vp),
    ]


class RefineDesc(C.Structure):
    _fields_ = [
        ("d_spectrum", _vp), ("d_work", _vp), ("n_pix", C.c_int32), ("n_peaks", C.c_int32),
        ("d_peaks", _vp), ("d_n_peaks", _vp), ("window_width", C.c_int32), ("pad", C.c_int32),
        ("d_window", _vp), ("d_fit", _vp), ("d_status", _vp), ("d_refined", _vp), ("d_n_refined", _vp),
    ]


PEAK_STATUS = ("ok", "no_data", "not_finite", "fit_failed", "negative", "outside", "empty_window", "too_few")
REFINE_MAX_WINDOW = 32

STRUCTS = [HoughDesc, HoughStats, VoteDesc, RansacDesc, RansacResult, RefitResult, MatchDesc, MatchResult,
           PairsDesc, PrepareDesc, FindPeaksDesc, RefineDesc]
HYP_ABORT, HYP_INTERCEPT, HYP_MONOTONIC, HYP_EMPTY, HYP_SCORED, HYP_UNSUPPORTED = range(6)  # rb2_hyp_status
BASIS_CODES = {"poly": 0, "legendre": 1, "chebyshev": 2}
COUNTER_NAMES = ("drawn", "aborted", "intercept", "monotonic", "empty", "scored", "eligible", "unsupported")

# every symbol include/rascal_b200.h declares
EXPORTS = {
    "rb2_abi_version": (C.c_int, []),
    "rb2_status_string": (C.c_char_p, [C.c_int]),
    "rb2_last_cuda_error": (C.c_char_p, []),
    "rb2_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rb2_struct_size": (C.c_int, [C.c_int]),
    "rb2_dfma_burn": (C.c_int, [C.c_int, C.c_int, _vp, _vp]),
    "rb2_hough_accumulate": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_int, _vp]),
    "rb2_hough_points": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp]),
    "rb2_candidate_vote": (C.c_int, [C.c_int, _vp, _vp, _vp]),
    "rb2_candidate_points": (C.c_int, [_vp, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp]),
    "rb2_ransac_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "rb2_ransac_chunk_plan": (C.c_int, [C.c_int, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rb2_ransac_batch": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "rb2_merge_winners": (C.c_int, [C.c_int, C.c_int, _vp, _vp, _vp]),
    "rb2_robust_polyfit": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, C.c_int, C.c_int, _vp, _vp]),
    "rb2_adjust_polyfit": (C.c_int, [C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int,
                                     C.c_double, C.c_double, _vp, _vp]),
    "rb2_hough_brute_force": (C.c_int, [_vp, _vp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _vp, _vp, _vp, _vp]),
    "rb2_hough_bin_points": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp]),
    "rb2_expand_pairs": (C.c_int, [C.c_int, _vp, _vp, _vp]),
    "rb2_ransac_prepare": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp]),
    "rb2_match_peaks": (C.c_int, [C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "rb2_refit_winners": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_double, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "rb2_find_peaks": (C.c_int, [C.c_int, _vp, _vp, _vp]),
    "rb2_refine_scratch_bytes": (C.c_size_t, [C.c_int]),
    "rb2_refine_peaks": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and check its ABI; raises if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing: build it with `python -m rascal_b200.csrc.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.rb2_abi_version() != ABI_VERSION:
        raise NativeError("ABI version mismatch")
    for i, s in enumerate(STRUCTS):
        if lib.rb2_struct_size(i) != C.sizeof(s):
            raise NativeError(f"struct layout mismatch for {s.__name__}: "
                              f"C {lib.rb2_struct_size(i)} vs ctypes {C.sizeof(s)}")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        lib = load()
        msg = lib.rb2_status_string(rc).decode()
        if rc == -1:
            msg += ": " + lib.rb2_last_cuda_error().decode()
        raise NativeError(f"{what} failed: {msg}")


def struct_array(cls, n):
    return (cls * n)()


def as_numpy_bytes(arr):
    """View a ctypes struct array as uint8 numpy (for upload)."""
    return np.frombuffer(arr, dtype=np.uint8)
