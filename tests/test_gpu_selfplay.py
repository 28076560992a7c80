"""Playing strength: the engine as player 0 against random players wins as often as the unmodified
reference does (tests/golden/ref_strength.json: SOSolver<Card,Sequential>{1000}, 200 games), at the same
budget (1000 iterations, one tree).  This is the end-to-end statistical parity of move selection."""
import json
import os

import numpy as np
import pytest

from ismcsolver_b200 import capi
from ismcsolver_b200.selfplay import play_games

pytestmark = pytest.mark.gpu
GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_strength.json")))


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=lambda c: f"kw{c['players']}")
def test_win_rate_against_random_matches_the_reference(engine, case):
    games = 400
    rep = play_games(engine, games, players=case["players"], solver_seats=(0,), iterations=case["iterations"],
                     trees=1, seed=77)
    assert sum(rep.wins) == games
    p_ref = case["wins"][0] / case["games"]
    p = rep.wins[0] / games
    se = np.sqrt(p_ref * (1 - p_ref) / case["games"] + p * (1 - p) / games)
    # stated tolerance: 4 combined standard errors
    assert abs(p - p_ref) <= 4 * se, (p, p_ref, se)
    # the random seats are interchangeable: none of them may beat the solver seat
    assert rep.wins[0] > max(rep.wins[1:])


def test_self_play_all_seats_and_time_mode(engine):
    """Every seat searches (benchmark.cpp:153-160 'versus self'), decisions limited by time."""
    rep = play_games(engine, 32, players=4, solver_seats=(0, 1, 2, 3), iterations=0, time_ns=2_000_000, trees=64, seed=5)
    assert sum(rep.wins) == 32 and rep.decisions == rep.plies
    assert rep.iterations > 0
