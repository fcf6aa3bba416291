"""CPU-side checks of the C-ABI library: it loads, exports every symbol
include/rascal_b200.h declares, and the ctypes struct layouts match.  No
compute call is made (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from rascal_b200.csrc import build
    from rascal_b200 import _native

    build.build()
    return _native.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "rascal_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rb2_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_declare_the_same_symbols():
    from rascal_b200 import _native

    assert header_functions() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    for name in header_functions():
        assert hasattr(lib, name), name


def test_integration_guide_names_every_entry_point():
    """INTEGRATION.md shows, per entry point, the reference code it replaces: none may be missing."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert [f for f in header_functions() if f not in doc] == []


def test_struct_layouts(lib):
    from rascal_b200 import _native

    for i, s in enumerate(_native.STRUCTS):
        assert lib.rb2_struct_size(i) == C.sizeof(s)
    assert lib.rb2_struct_size(99) == -1


def test_diagnostics(lib):
    assert lib.rb2_abi_version() == 2
    assert lib.rb2_status_string(0) == b"ok"
    assert lib.rb2_status_string(-2) == b"bad argument"


def test_argument_validation_without_device(lib):
    # argument errors are reported before any CUDA work
    assert lib.rb2_hough_accumulate(0, None, None, None, 0, None) == -2
    assert lib.rb2_candidate_vote(1, None, None, None) == -2
    assert lib.rb2_ransac_batch(1, None, None, None, None, 0, None) == -2
    assert lib.rb2_refit_winners(1, None, None, None, 0.0, 0, None, None, None, None, 1, None) == -2


def test_chunk_plan_covers_every_call(lib):
    """rb2_ransac_chunk_plan is host logic (no device): the chunks cover the call, one chunk never exceeds the queue
    capacity of 2^25 entries over all problems, small calls are not cut below 2^23 ids in total."""
    pp, nc, nl = C.c_longlong(0), C.c_int(0), C.c_int(0)
    for n_problems, max_hyp in [(1, 100_000_000), (1, 12_500_000), (1, 1), (1, 0), (2048, 10_000), (7, 3_000_001),
                                (1 << 20, 5), (1, (1 << 40) - 1)]:
        assert lib.rb2_ransac_chunk_plan(n_problems, max_hyp, C.byref(pp), C.byref(nc), C.byref(nl)) == 0
        assert pp.value >= 1 and pp.value * n_problems <= max(1 << 25, n_problems)
        assert nc.value * pp.value >= max_hyp and (nc.value - 1) * pp.value < max(max_hyp, 1)
        assert 1 <= nl.value <= 4 and nl.value <= max(nc.value, 1)
        if nc.value > 1:
            assert pp.value * n_problems > (1 << 23) - n_problems or pp.value == (1 << 25) // n_problems
    assert lib.rb2_ransac_chunk_plan(0, 10, None, None, None) == -2
    assert lib.rb2_ransac_chunk_plan(1, -1, None, None, None) == -2


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises instead of computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import numpy as np

    from rascal_b200 import _native, engine

    with pytest.raises(_native.NativeError):
        engine.HoughBatch([dict(pair_x=np.zeros(1), pair_y=np.zeros(1), min_slope=0.0, max_slope=1.0,
                                min_intercept=0.0, max_intercept=1.0, n_slopes=2, xbins=2, ybins=2)])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rascal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
