"""Host-side behaviour of rascal_b200.util / batch.peaks_from_spectra that needs no GPU:
argument checking happens before any device call."""
import numpy as np
import pytest


def test_find_peaks_rejects_what_the_device_path_does_not_implement():
    from rascal_b200 import util

    x = np.sin(np.arange(50) / 3.0)
    for kw in (dict(threshold=0.1), dict(width=2), dict(wlen=11), dict(rel_height=0.5), dict(plateau_size=1)):
        with pytest.raises(NotImplementedError):
            util.find_peaks(x, **kw)
    with pytest.raises(NotImplementedError):
        util._scalar_min("height", (0.0, 1.0))  # SciPy's (min, max) interval form
    with pytest.raises(TypeError):
        util.find_peaks(x, bogus=1)
    with pytest.raises(ValueError):
        util.find_peaks(np.zeros((3, 3)))
    assert np.isnan(util._scalar_min("height", None)) and util._scalar_min("distance", 3) == 3.0


def test_peaks_from_spectra_leaves_entries_with_peaks_alone():
    from rascal_b200 import batch

    specs = [dict(peaks=np.array([3.0, 1.0]), lines=np.array([1.0])), dict(peaks=[5.0], spectrum=np.zeros(10), lines=[2.0])]
    out = batch.peaks_from_spectra(specs)
    assert out[0] is specs[0] and out[1] is specs[1]


def test_oracle_is_not_imported_by_the_product():
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rascal_b200")
    for name in os.listdir(root):
        if name.endswith(".py"):
            src = open(os.path.join(root, name)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), name
