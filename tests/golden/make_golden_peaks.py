#!/usr/bin/env python
"""Golden fixtures for the step before the hot path: the UNMODIFIED
reference's ``rascal.util.refine_peaks`` and ``scipy.signal.find_peaks`` (the recipe of every
reference example / test: test/test_xe_calibration.py:69-70, examples/*.py) on three real arcs
shipped with the reference and on synthetic arcs with blends, clipped windows and flat tops.

    python tests/golden/make_golden_peaks.py     # build container only (needs /root/reference)

Each tests/golden/peaks_<case>.npz holds: spectrum, the find_peaks arguments and its result
(peaks, prominences where asked), window_width and the refined peaks.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402,F401  (installs the shims, imports the reference)
from make_golden_configs import EX, read_fits_image  # noqa: E402
from rascal import util  # noqa: E402
from scipy.signal import find_peaks  # noqa: E402


def synthetic_arc(seed, n_pix=1024, n_lines=60, noise=3.0, sat=None, blend=True):
    rng = np.random.default_rng(seed)
    x = np.arange(n_pix, dtype=np.float64)
    centres = np.sort(rng.uniform(2, n_pix - 2, n_lines))
    if blend:
        centres[5] = centres[4] + 3.3  # overlapping refinement windows and a blended pair
        centres[20] = centres[19] + 7.1
    amps = rng.uniform(200, 20000, n_lines)
    sig = rng.uniform(0.9, 1.8, n_lines)
    s = 50.0 + 0.01 * x
    for c, a, w in zip(centres, amps, sig):
        s += a * np.exp(-0.5 * ((x - c) / w) ** 2)
    s += rng.normal(0, noise, n_pix)
    if sat is not None:
        s = np.minimum(s, sat)  # flat-topped (saturated) lines
    return s


def record(name, spectrum, fp_kw, window_width):
    spectrum = np.asarray(spectrum, dtype=np.float64)
    peaks, props = find_peaks(spectrum, **fp_kw)
    refined = util.refine_peaks(spectrum.copy(), peaks, window_width=window_width)
    out = dict(spectrum=spectrum, peaks=peaks.astype(np.int64), refined=np.asarray(refined, float),
               window_width=np.int64(window_width))
    for k in ("height", "distance", "prominence"):
        out["fp_" + k] = np.float64(fp_kw[k]) if fp_kw.get(k) is not None else np.float64(np.nan)
    if "prominences" in props:
        out["prominences"] = props["prominences"]
    np.savez_compressed(os.path.join(HERE, f"peaks_{name}.npz"), **out)
    print(f"{name}: {len(spectrum)} px, {len(peaks)} peaks -> {len(refined)} refined")


def main():
    data, _ = read_fits_image(os.path.join(EX, "data_lt_sprat/v_a_20190516_57_1_0_1.fits"))
    record("sprat_xe", np.median(data[110:120], axis=0), dict(height=300, prominence=200, distance=5), 5)
    spec = np.array(json.load(open(os.path.join(EX, "data_keck_deimos/keck_deimos_830g_l_PYPIT.json")))["spec"])
    record("deimos", spec, dict(prominence=200, distance=10), 3)
    d = np.loadtxt("/root/reference/test/ground_truth/gmos-effective-pixel-spectrum.txt", delimiter=",")
    record("gmos", d[:, 1], dict(height=1000, prominence=500, distance=5), 3)
    record("synth_a", synthetic_arc(1), dict(height=100, distance=5), 5)
    record("synth_b", synthetic_arc(2, n_pix=2048, n_lines=90, noise=8.0), dict(prominence=60, distance=3), 10)
    record("synth_sat", synthetic_arc(3, sat=12000.0), dict(height=150, prominence=100, distance=4), 4)
    record("synth_edge", synthetic_arc(4, n_pix=300, n_lines=25, blend=False), dict(height=80), 10)


if __name__ == "__main__":
    main()
