#!/usr/bin/env python
"""Record HoughTransform.generate_hough_points_brute_force + bin_hough_points (+ add_hough_points)
of the UNMODIFIED reference (houghtransform.py:94-210) on a subset of each golden case's pairs.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_brute.py        -> tests/golden/brute_<case>.npz
"""
import hashlib
import os
import sys
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore")
sys.modules.setdefault("pynverse", types.ModuleType("pynverse"))
sys.path.insert(0, "/root/reference/src")

from rascal.houghtransform import HoughTransform  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def canon(points):
    """Row order of the reference depends on np.argsort's (unstable) tie order: compare as a sorted multiset."""
    p = np.asarray(points)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def main():
    for case, stride in (("c3_deg4", 23), ("poly_quadratic_deg2", 3)):
        g = np.load(os.path.join(HERE, case + ".npz"), allow_pickle=True)
        pairs = g["pairs"][::stride]
        lo, hi, mi, ma = g["limits"]
        ht = HoughTransform()
        ht.set_constraints(lo, hi, mi, ma)
        ht.generate_hough_points_brute_force(pairs[:, 0], pairs[:, 1])
        pts = np.array(ht.hough_points)
        ht.bin_hough_points(64, 48)
        out = dict(pairs=pairs, limits=g["limits"], n_points=len(pts),
                   points_sorted_sha256=hashlib.sha256(np.ascontiguousarray(canon(pts)).tobytes()).hexdigest(),
                   hist=ht.hist, xedges=ht.xedges, yedges=ht.yedges)
        # add_hough_points: extend with a shifted copy of the first 1000 points, re-bin
        extra = pts[:1000] + np.array([1e-3, 5.0])
        ht.add_hough_points(extra)
        ht.bin_hough_points(40, 40)
        out.update(extra=extra, hist_added=ht.hist, xedges_added=ht.xedges, yedges_added=ht.yedges)
        np.savez_compressed(os.path.join(HERE, "brute_" + case + ".npz"), **out)
        print(case, len(pairs), "pairs ->", len(pts), "points")


if __name__ == "__main__":
    main()
