#!/usr/bin/env python
"""Golden fixtures for BASELINE configs[3] (C4, the 10k-spectrum batch): three spectra of the batch - one
per detector size 1024 / 2048 / 4096 - recorded from the UNMODIFIED reference with make_golden.py's hooks,
so that the many-spectra path has an oracle on its own inputs (ragged peak counts, perturbed priors,
spurious peaks, distractor lines).

    python tests/golden/make_golden_c4.py     # build container only (needs /root/reference)

The inputs are bench.c4_input(i) (SURVEY.md 8d "C4 concrete input"); the index of each recorded spectrum
is stored in the fixture's config.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as G  # noqa: E402  (installs the shims, imports the reference)
from bench import c4_input  # noqa: E402


def main():
    want = {1024: None, 2048: None, 4096: None}
    i = 0
    while any(v is None for v in want.values()):
        sp = c4_input(i)
        if want.get(sp["num_pix"], 0) is None:
            want[sp["num_pix"]] = i
        i += 1
    for npix, idx in sorted(want.items()):
        sp = c4_input(idx)
        G.record_case("c4_npix%d" % npix, sp["peaks"], sp["lines"], spectrum=np.zeros(npix),
                      hough=sp["hough"], ransac=sp["ransac"], fit=dict(max_tries=80, **sp["fit"]),
                      seed=10_000 + idx, atlas_kw=dict(candidate_tolerance=10.0))
        # remember which spectrum of the batch this is
        path = os.path.join(HERE, "c4_npix%d.npz" % npix)
        z = dict(np.load(path, allow_pickle=False))
        z["c4_index"] = np.int32(idx)
        np.savez_compressed(path, **z)


if __name__ == "__main__":
    main()
