"""The oracle against the recordings of the random problems tests/live_problems.py::LIVE_CASES (unweighted vote,
filter_close, over-determined samples, hough_weight != 1, Legendre / Chebyshev at degree 3 / 4): the same bit-exact
bars as tests/test_oracle_golden.py.  The fixtures exist so that the DEVICE path meets these problems on the GPU
box (tests/test_gpu_live_cases.py); tests/test_oracle_live_reference.py draws further ones against the live
reference where its checkout exists."""
import pytest

from tests import test_oracle_golden as T
from tests.conftest import load_golden
from tests.live_problems import KNOWN_CASES, LIVE_CASES


@pytest.fixture(scope="module", params=["live_%d" % s for s in LIVE_CASES] + ["live_known_%d" % s for s in KNOWN_CASES])
def live(request):
    return load_golden(request.param)


def test_hough_and_candidates(live):
    T.test_limits_and_pairs(live)
    T.test_hough_points_and_histogram(live)
    T.test_reference_order_is_a_valid_ranking(live)
    T.test_candidates(live)


def test_draws_loop_and_final_model(live, request):
    T.test_replayed_hypotheses(live)
    T.test_seeded_loop_reproduces_reference(live, request)
    T.test_order_independent_rule(live)


def test_host_setup_tables(live):
    """engine.RansacSetup (the host side of K3's set-up, calibrator.py:441-523) on the same problems: peaks,
    candidate lists, sampling law, monotonicity range and the Hough-weight spline tables against the oracle."""
    from tests import test_host_logic as H

    H.test_ransac_setup_tables(live)
    H.test_weight_tables_reproduce_the_reference_spline(live)
