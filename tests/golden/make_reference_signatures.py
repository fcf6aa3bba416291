#!/usr/bin/env python
"""tests/golden/reference_signatures.json: names, order and defaults of the arguments of every public method of the
reference's Calibrator / HoughTransform / util functions that the drop-in mirrors, read with
inspect.signature from the UNMODIFIED reference.

    python tests/golden/make_reference_signatures.py     # build container only (needs /root/reference)
"""
import inspect
import json
import os
import sys
import types

sys.modules.setdefault("pynverse", types.ModuleType("pynverse"))
sys.path.insert(0, "/root/reference/src")

from rascal import util  # noqa: E402
from rascal.calibrator import Calibrator  # noqa: E402
from rascal.houghtransform import HoughTransform  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CALIBRATOR = ["__init__", "set_calibrator_properties", "set_hough_properties", "set_ransac_properties", "set_atlas",
              "set_known_pairs", "do_hough_transform", "save_hough_transform", "load_hough_transform", "fit", "match_peaks"]
HOUGH = ["set_constraints", "generate_hough_points", "generate_hough_points_brute_force", "add_hough_points",
         "bin_hough_points", "save", "load"]
UTIL = ["refine_peaks"]


def sig(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        if name == "self":
            continue
        d = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([name, d])
    return out


def main():
    doc = {"Calibrator": {m: sig(getattr(Calibrator, m)) for m in CALIBRATOR},
           "HoughTransform": {m: sig(getattr(HoughTransform, m)) for m in HOUGH},
           "util": {m: sig(getattr(util, m)) for m in UTIL}}
    with open(os.path.join(HERE, "reference_signatures.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    print(json.dumps(doc)[:400])


if __name__ == "__main__":
    main()
