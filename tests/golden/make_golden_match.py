#!/usr/bin/env python
"""Record Calibrator.match_peaks of the UNMODIFIED reference on the golden cases.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_match.py

For every fixture of make_golden.py the reference Calibrator is given the same
peaks and atlas lines and `match_peaks(fit_coeff=<recorded fit>, tolerance=t)` is
called for a list of tolerances, with robust_refit True.  The reference raises
as soon as one peak has two atlas lines inside the tolerance (TypeError at
calibrator.py:1867 under this numpy); the exception type is recorded for those
tolerances and the 7-tuple for the others -> tests/golden/match_<case>.npz.
A second atlas per case, thinned to lines with no neighbour inside 12 A (what
the reference's own Atlas(min_atlas_distance=...) separation filter produces),
gives the reference a defined run with many matches at the default tolerances:
keys with the suffix "_sep".
"""
import json
import os
import sys
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.modules.setdefault("pynverse", types.ModuleType("pynverse"))
sys.path.insert(0, "/root/reference/src")

from rascal.atlas import Atlas  # noqa: E402
from rascal.calibrator import Calibrator  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TOLERANCES = (10.0, 5.0, 2.0, 1.0, 0.5, 0.1)


def main():
    for case in (sys.argv[1:] or ("c3_deg4", "fine_deg5", "poly_linear_deg1", "poly_quadratic_deg2", "c1_sprat_xe", "c5_deimos", "c2_gmos_cuar")):
        g = np.load(os.path.join(HERE, case + ".npz"), allow_pickle=True)
        peaks, lines, coeff = g["peaks"], g["lines"], g["fit_coeff"]
        c = Calibrator(peaks, spectrum=np.zeros(len(g["pixel_list"])))
        a = Atlas()
        a.add_user_atlas(["X"] * len(lines), list(lines))
        c.atlas = a
        c.polyfit = np.polynomial.polynomial.polyfit   # what fit(fit_type='poly') installs, :1621-1623
        c.polyval = np.polynomial.polynomial.polyval
        out = dict(peaks=peaks, lines=np.asarray(a.get_lines()), fit_coeff_in=coeff)
        sl = np.sort(lines)
        gap = np.diff(sl)
        lonely = np.concatenate(([True], gap > 12.0)) & np.concatenate((gap > 12.0, [True]))
        a_sep = Atlas()
        a_sep.add_user_atlas(["X"] * int(lonely.sum()), list(sl[lonely]))
        out["lines_sep"] = np.asarray(a_sep.get_lines())
        status = {}
        for t, atlas, sfx in [(t, a, "") for t in TOLERANCES] + [(t, a_sep, "_sep") for t in (5.0, 2.0)]:
            key = ("%g" % t).replace(".", "p") + sfx
            c.atlas = atlas
            try:
                r = c.match_peaks(fit_coeff=coeff.copy(), tolerance=t)
            except Exception as e:  # ambiguous peak
                status[key] = type(e).__name__
                continue
            status[key] = "ok"
            out["fit_coeff_" + key] = np.asarray(r[0])
            out["matched_peaks_" + key] = np.asarray(r[1])
            out["matched_atlas_" + key] = np.asarray(r[2])
            out["rms_" + key] = r[3]
            out["residuals_" + key] = np.asarray(r[4])
            out["utilisation_" + key] = np.array([r[5], r[6]])
        out["status_json"] = json.dumps(status)
        np.savez_compressed(os.path.join(HERE, "match_" + case + ".npz"), **out)
        print(case, status)


if __name__ == "__main__":
    main()
