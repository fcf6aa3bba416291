"""Shared helpers for the parity tests (oracle side only)."""
import numpy as np

from oracle import rascal_oracle as O


def oracle_from_golden(g, refit=O.robust_polyfit, max_tries=None, trace=None):
    """Run the oracle's whole path on a golden case's inputs and settings."""
    cfg = g["config"]
    kw = {}
    kw.update(cfg["hough"])
    kw.update(cfg["ransac"])
    assert kw.pop("linear", True)  # the linear candidate search is the path restated here
    kw.update(cfg["fit"])
    kw.pop("fit_tolerance", None)  # only a warning threshold in Calibrator.fit (:1696)
    if max_tries is not None:
        kw["max_tries"] = max_tries
    if cfg.get("known"):
        kw["pix_known"], kw["wave_known"] = (np.array(v, dtype=float) for v in cfg["known"])
    return O.run_path(g["peaks"], g["lines"], pixel_list=g["pixel_list"],
                      seed=cfg["seed"], refit=refit, trace=trace, **kw)


def problem_from_golden(g):
    """RansacProblem built from the golden candidates + golden histogram."""
    cfg = g["config"]
    r, f = cfg["ransac"], cfg["fit"]
    lo, hi, min_i, max_i = g["limits"]
    return O.RansacProblem(
        g["candidate_peak"], g["candidate_arc"], fit_deg=f.get("fit_deg", 4),
        sample_size=int(g["sample_size"]),
        ransac_tolerance=r.get("ransac_tolerance", 5),
        min_intercept=min_i, max_intercept=max_i, pixel_list=g["pixel_list"],
        hist=g["hist"].astype(float), xedges=g["xedges"], yedges=g["yedges"],
        hough_weight=r.get("hough_weight", 1.0),
        filter_close=r.get("filter_close", False),
        minimum_matches=min(r.get("minimum_matches", 0), len(g["lines"]),
                            len(g["peaks"])),
        minimum_peak_utilisation=r.get("minimum_peak_utilisation", 0.0),
        minimum_fit_error=r.get("minimum_fit_error", 1e-4),
        n_peaks_total=len(g["peaks"]), fit_type=f.get("fit_type", "poly"),
        pix_known=None if not cfg.get("known") else np.array(cfg["known"][0], dtype=float),
        wave_known=None if not cfg.get("known") else np.array(cfg["known"][1], dtype=float))
