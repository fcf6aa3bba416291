#!/usr/bin/env python
"""Golden fixtures for the BASELINE configs that use REAL data (C1 LT/SPRAT Xe, C2 GMOS CuAr
ground-truth spectrum, C5 Keck/DEIMOS) from the UNMODIFIED reference, with the same recording
hooks as make_golden.py (which this script imports).  Peaks come from the reference's own
recipe (scipy.signal.find_peaks + rascal.util.refine_peaks, examples/*.py), atlas lines from
rascal.atlas.Atlas (NIST list shipped with the reference).

    python tests/golden/make_golden_configs.py     # build container only (needs /root/reference)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (installs the shims, imports the reference)
from rascal import util  # noqa: E402
from rascal.atlas import Atlas  # noqa: E402
from scipy.signal import find_peaks  # noqa: E402

EX = "/root/reference/examples"


def read_fits_image(path):
    """Primary HDU of a simple FITS file (astropy is not installed here)."""
    with open(path, "rb") as f:
        raw = f.read()
    hdr, off = {}, 0
    while True:
        block = raw[off:off + 2880].decode("ascii", "replace")
        off += 2880
        for i in range(0, 2880, 80):
            card = block[i:i + 80]
            key = card[:8].strip()
            if key == "END":
                break
            if card[8:10] == "= ":
                hdr[key] = card[10:].split("/")[0].strip().strip("'").strip()
        else:
            continue
        break
    bitpix, nx, ny = int(hdr["BITPIX"]), int(hdr["NAXIS1"]), int(hdr["NAXIS2"])
    dt = {16: ">i2", 32: ">i4", -32: ">f4", -64: ">f8"}[bitpix]
    data = np.frombuffer(raw, dtype=dt, count=nx * ny, offset=off).reshape(ny, nx).astype(np.float64)
    data = data * float(hdr.get("BSCALE", 1.0)) + float(hdr.get("BZERO", 0.0))
    return data, hdr


def c1_sprat():
    """examples/example_lt_sprat_xe.py:12-67."""
    data, hdr = read_fits_image(os.path.join(EX, "data_lt_sprat/v_a_20190516_57_1_0_1.fits"))
    spectrum = np.median(data[110:120], axis=0)
    peaks, _ = find_peaks(spectrum, height=300, prominence=200, distance=5)
    peaks = util.refine_peaks(spectrum.copy(), peaks, window_width=5)
    atlas = Atlas(elements=["Xe"], min_intensity=100.0, min_distance=10, min_atlas_wavelength=3500.0,
                  max_atlas_wavelength=8000.0, range_tolerance=500.0, pressure=float(hdr["REFPRES"]) * 100.0,
                  temperature=float(hdr["REFTEMP"]), relative_humidity=float(hdr["REFHUMID"]))
    G.record_case("c1_sprat_xe", np.asarray(peaks, float), np.asarray(atlas.get_lines(), float),
                  spectrum=np.zeros(len(spectrum)),
                  hough=dict(num_slopes=5000, range_tolerance=500.0, xbins=200, ybins=200,
                             min_wavelength=3500.0, max_wavelength=8000.0),
                  ransac=dict(sample_size=5, top_n_candidate=5, filter_close=True),
                  fit=dict(max_tries=500, candidate_tolerance=2.0), seed=0, atlas_kw=dict(candidate_tolerance=5.0))


def c5_deimos():
    """examples/example_keck_deimos_830g.py:10-61 with explicit atlas limits and the finer
    Hough binning of BASELINE configs[4], reduced to 400 x 400 to keep the fixture small."""
    spec = np.array(json.load(open(os.path.join(EX, "data_keck_deimos/keck_deimos_830g_l_PYPIT.json")))["spec"])
    peaks, _ = find_peaks(spec, prominence=200, distance=10)
    peaks = util.refine_peaks(spec.copy(), peaks, window_width=3)
    atlas = Atlas(elements=["Ne", "Ar", "Kr"], min_intensity=1000.0, pressure=70000.0, temperature=285.0,
                  range_tolerance=500.0, min_atlas_wavelength=6500.0, max_atlas_wavelength=10500.0)
    G.record_case("c5_deimos", np.asarray(peaks, float), np.asarray(atlas.get_lines(), float),
                  spectrum=np.zeros(len(spec)),
                  hough=dict(num_slopes=2000, range_tolerance=500.0, linearity_tolerance=50, xbins=400, ybins=400,
                             min_wavelength=6500.0, max_wavelength=10500.0),
                  ransac=dict(sample_size=5, top_n_candidate=5, linear=True, filter_close=True, ransac_tolerance=5,
                              candidate_weighted=True, hough_weight=1.0),
                  fit=dict(max_tries=300), seed=0)


def c5_deimos_full():
    """BASELINE configs[4] at its stated size: num_slopes x bins = 2000 x (1000 x 1000) - 10^6 Hough lines, which
    costs the reference ~1-2 minutes of candidate collection (calibrator.py:219-261).  Few tries: this fixture
    pins K1 (histogram, edges), the ranking and K2 (candidates) at 10^6 bins; the RANSAC trace stays short."""
    spec = np.array(json.load(open(os.path.join(EX, "data_keck_deimos/keck_deimos_830g_l_PYPIT.json")))["spec"])
    peaks, _ = find_peaks(spec, prominence=200, distance=10)
    peaks = util.refine_peaks(spec.copy(), peaks, window_width=3)
    atlas = Atlas(elements=["Ne", "Ar", "Kr"], min_intensity=1000.0, pressure=70000.0, temperature=285.0,
                  range_tolerance=500.0, min_atlas_wavelength=6500.0, max_atlas_wavelength=10500.0)
    G.record_case("c5_deimos_full", np.asarray(peaks, float), np.asarray(atlas.get_lines(), float),
                  spectrum=np.zeros(len(spec)),
                  hough=dict(num_slopes=2000, range_tolerance=500.0, linearity_tolerance=50, xbins=1000, ybins=1000,
                             min_wavelength=6500.0, max_wavelength=10500.0),
                  ransac=dict(sample_size=5, top_n_candidate=5, linear=True, filter_close=True, ransac_tolerance=5,
                              candidate_weighted=True, hough_weight=1.0),
                  fit=dict(max_tries=40), seed=0)


def c2_gmos():
    """BASELINE configs[1]: Gemini GMOS longslit CuAr arc, degree 5.  The example's FITS frame is not
    shipped; the reference's own ground-truth effective-pixel spectrum of the same
    instrument is (test/ground_truth/gmos-effective-pixel-spectrum.txt), with the example's settings
    (examples/example_gemini_gmos_longslit.py:37-61) and fit_deg=5."""
    d = np.loadtxt("/root/reference/test/ground_truth/gmos-effective-pixel-spectrum.txt", delimiter=",")
    spectrum = d[:, 1]
    peaks, _ = find_peaks(spectrum, height=1000, prominence=500, distance=5)
    peaks = util.refine_peaks(spectrum.copy(), peaks, window_width=3)
    atlas = Atlas(elements=["Cu", "Ar"], min_intensity=80, min_distance=5, range_tolerance=500.0, pressure=70000.0,
                  temperature=280.0, min_atlas_wavelength=5000.0, max_atlas_wavelength=9500.0)
    G.record_case("c2_gmos_cuar", np.asarray(peaks, float), np.asarray(atlas.get_lines(), float),
                  spectrum=np.zeros(len(spectrum)),
                  hough=dict(num_slopes=5000, range_tolerance=500.0, xbins=200, ybins=200,
                             min_wavelength=5000.0, max_wavelength=9500.0),
                  ransac=dict(sample_size=6, top_n_candidate=5),
                  fit=dict(max_tries=300, fit_deg=5), seed=0)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c5"]
    if "c2" in which:
        c2_gmos()
    if "c1" in which:
        c1_sprat()
    if "c5" in which:
        c5_deimos()
    if "c5full" in which:
        c5_deimos_full()
