"""The device path on the recorded random problems of tests/live_problems.py::LIVE_CASES - configurations the
other goldens are thin on (candidate_weighted=False, filter_close, over-determined samples, hough_weight != 1,
Legendre / Chebyshev at degree 3 / 4) - with the bars of the golden tests: K1 histogram / edges / ranking / points
bit-exact, K2 candidates identical, every recorded draw replayed through K3 with identical decisions and counts,
the reference's winner, K4's model."""
import pytest

from tests import test_gpu_hough_vote as HV
from tests import test_gpu_ransac_refit as RR
from tests.conftest import load_golden
from tests.live_problems import KNOWN_CASES, LIVE_CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=LIVE_CASES)
def live(request):
    return load_golden("live_%d" % request.param)


def test_hough(live):
    HV.test_hough_golden(live)
    HV.test_hough_points_golden(live)


def test_vote(live):
    HV.test_vote_golden(live)


def test_setup_and_replayed_draws(live):
    RR.test_setup_matches_reference_tables(live)
    RR.test_replay_reference_draws(live)
    RR.test_replay_with_reference_coefficients_is_decision_exact(live)


def test_winner_and_refit(live, request):
    RR.test_replay_selects_reference_winner(live, request)


def test_shared_memory_opt_in_survives_mixed_calls():
    """One process, one degree: an exactly determined fit with large tables, an over-determined (least-squares) fit
    with small ones, then an exactly determined fit in between.  The match kernel is shared by both instantiations of
    the chunk launcher; the second call used to LOWER its dynamic-shared-memory cap below what the third call asks
    for (rb2_ransac_batch: "invalid argument") - found by these random problems.  The cap is now kept per kernel,
    grow-only, never below 48 KB."""
    from rascal_b200 import engine

    for case in ("c3_deg4", "live_20264", "live_30003", "live_20264", "c3_deg4",       # degree 4: exact, LSQ, exact ...
                 "live_20262", "live_30010", "live_31007", "live_30002"):             # degree 3
        g = load_golden(case)
        s = RR.setup_from_golden(g)
        rb = engine.RansacBatch([s])
        n = min(len(g["tr_status"]), 2000)
        out = rb.run(0, n, replay=(g["tr_x_idx"][:n], g["tr_y_idx"][:n]), per_hypothesis=("status",))
        assert (out["status"] == g["tr_status"][:n].astype("uint8")).all(), case


@pytest.mark.parametrize("seed", KNOWN_CASES)
def test_known_pairs(seed, request):
    """set_known_pairs (calibrator.py:1507-1550, :568-573): the user's (pixel, wavelength) pairs are appended to every
    sample, so every hypothesis is a least-squares fit over sample + known rows (rb2_ransac_desc.n_known).  Recorded
    from the reference with known pairs set; replayed draw for draw, winner and K4's model as for every golden."""
    g = load_golden("live_known_%d" % seed)
    assert g["config"]["known"]
    RR.test_setup_matches_reference_tables(g)
    RR.test_replay_reference_draws(g)
    RR.test_replay_with_reference_coefficients_is_decision_exact(g)
    RR.test_replay_selects_reference_winner(g, request)
