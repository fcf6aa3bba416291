#!/usr/bin/env python
"""Install the UNMODIFIED reference (rascal v0.3.9, pure Python) for the timed CPU arm of bench.py.

    python scripts/make_baseline_ref.py        # build container only: reads /root/reference

`pip install /root/reference` cannot be used offline (setup.cfg pins numpy < 1.24 and pulls pynverse, which is
not in the wheelhouse), so the package directory is copied into the git-ignored baseline/_ref/ - which
travels to the GPU box with the snapshot - with the four numpy-2 compatibility patches of SURVEY.md App. B,
none of which changes the arithmetic of the path:

  calibrator.py   `self.pairs == []` -> `len(self.pairs) == 0` (two places; the comparison raises for arrays)
  calibrator.py   `match[j]` -> `match[int(j)]` in match_peaks (array index)
  houghtransform.py  np.save of a ragged list needs dtype=object
  pynverse        a stub module (only SyntheticSpectrum.get_pixels for degree >= 2 uses it)

Nothing under baseline/_ref is imported by the product; bench.py's `--impl reference` arm and its
`cpu_baseline` leg are the only users.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/rascal"
DST = os.path.join(ROOT, "baseline", "_ref")


def patch(path, edits):
    s = open(path).read()
    for old, new, count in edits:
        assert s.count(old) == count, "%s: expected %d x %r, found %d" % (path, count, old, s.count(old))
        s = s.replace(old, new)
    open(path, "w").write(s)


def main():
    if not os.path.isdir(SRC):
        print("no reference checkout at %s: nothing to do" % SRC)
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(SRC, os.path.join(DST, "rascal"))
    patch(os.path.join(DST, "rascal", "calibrator.py"),
          [("self.pairs == []", "len(self.pairs) == 0", 2), ("match[j]", "match[int(j)]", 1)])
    hp = os.path.join(DST, "rascal", "houghtransform.py")
    s = open(hp).read()
    if "np.save(filename + \".npy\", output_npy)" in s:
        s = s.replace("np.save(filename + \".npy\", output_npy)",
                      "np.save(filename + \".npy\", np.array(output_npy, dtype=object))")
        open(hp, "w").write(s)
    with open(os.path.join(DST, "pynverse.py"), "w") as f:
        f.write('"""Stub: pynverse is not installed; only SyntheticSpectrum.get_pixels (degree >= 2) needs it."""\n\n\n'
                "def inversefunc(*a, **k):\n    raise NotImplementedError('pynverse is not available')\n")
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Copy of /root/reference/src/rascal (v0.3.9) made by scripts/make_baseline_ref.py with the numpy-2\n"
                "compatibility patches of SURVEY.md App. B.  Git-ignored; used only by bench.py's CPU arms.\n")
    print("installed", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
