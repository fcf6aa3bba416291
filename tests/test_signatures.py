"""Drop-in boundary: every public method the facade mirrors takes the reference's arguments - same
names, same order, same defaults (tests/golden/reference_signatures.json, read from the unmodified reference with
inspect.signature by make_reference_signatures.py); what the facade adds comes after them and is optional."""
import inspect
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import GOLDEN_DIR, ROOT

REF = json.load(open(os.path.join(GOLDEN_DIR, "reference_signatures.json")))


def _objects():
    from rascal_b200 import util
    from rascal_b200.calibrator import Calibrator, HoughTransform

    return {"Calibrator": Calibrator, "HoughTransform": HoughTransform, "util": util}


@pytest.mark.parametrize("owner,method", [(o, m) for o in sorted(REF) for m in sorted(REF[o])])
def test_same_arguments_and_defaults(owner, method):
    fn = getattr(_objects()[owner], method)
    got = [(n, p) for n, p in inspect.signature(fn).parameters.items() if n != "self"]
    want = REF[owner][method]
    assert len(got) >= len(want)
    for (name, p), (rname, rdefault) in zip(got, want):
        assert name == rname
        assert (None if p.default is inspect.Parameter.empty else repr(p.default)) == rdefault, (name, p.default, rdefault)
    for name, p in got[len(want):]:  # extensions (n_hypotheses, refit_mode, ...) never change a reference call
        assert p.default is not inspect.Parameter.empty, name


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/rascal"), reason="needs the reference checkout (build container only)")
def test_recorded_signatures_are_the_references(tmp_path):
    code = open(os.path.join(GOLDEN_DIR, "make_reference_signatures.py")).read().replace(
        "HERE = os.path.dirname(os.path.abspath(__file__))", "HERE = %r" % str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    assert json.load(open(os.path.join(tmp_path, "reference_signatures.json"))) == REF
