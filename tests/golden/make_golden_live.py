#!/usr/bin/env python
"""Fixtures live_<seed>.npz / live_known_<seed>.npz: the random problems tests/live_problems.py::LIVE_CASES and
KNOWN_CASES (the latter with set_known_pairs), solved by the UNMODIFIED
reference under the recording hooks of make_golden.py (same arrays as every other golden).

    python tests/golden/make_golden_live.py     # build container only (needs /root/reference)

One canonicalisation beside the bin ranking (make_golden.py, _CANON): with candidate_weighted=False the per-peak
aggregates are integer counts and the reference ranks them with NumPy's default unstable argsort
(calibrator.py:212), whose tie order depends on the NumPy build and the CPU; for those cases the reference's own
code runs with np.argsort(kind="stable") inside rascal.calibrator."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as G  # noqa: E402  (installs the shims, imports the reference)
import rascal.calibrator as rc  # noqa: E402
from tests.live_problems import KNOWN_CASES, LIVE_CASES, random_problem  # noqa: E402


class StableTies:
    def __getattr__(self, k):
        return getattr(np, k)

    @staticmethod
    def argsort(a, *args, **kw):
        kw.setdefault("kind", "stable")
        return np.argsort(a, *args, **kw)


def main():
    for seed, with_known in [(s, False) for s in LIVE_CASES] + [(s, True) for s in KNOWN_CASES]:
        spec = random_problem(seed, with_known=with_known)
        assert with_known == ("known" in spec)
        rc.np = np if spec["ransac"]["candidate_weighted"] else StableTies()
        try:
            G.record_case(("live_known_%d" % seed) if with_known else spec["name"], np.array(spec["peaks"]),
                          np.array(spec["lines"]), spectrum=np.zeros(spec["npix"]),
                          hough=spec["hough"], ransac=spec["ransac"], fit=spec["fit"], seed=spec["seed"],
                          atlas_kw=dict(candidate_tolerance=spec["fit"]["candidate_tolerance"]), known=spec.get("known"))
        finally:
            rc.np = np


if __name__ == "__main__":
    main()
