"""Host-side logic of the round-2 harnesses that needs no GPU: the agreement summary against golden reference
statistics, the synthetic position stream of bench.py, the geometry choices of bench.py, and the reference seat of
the self-play driver (when the reference harness has been built)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from ismcsolver_b200 import agreement as A
from ismcsolver_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_golden_1m_file_is_what_the_tests_expect():
    golden = A.load_golden()
    assert len(golden["cases"]) >= 20                                      # at least 20 positions
    views = [A.reference_view(c) for c in golden["cases"]]
    assert sum(v[0] for v in views) >= 15                                   # decisive positions
    for c, (decisive, modal, best_by_win, win) in zip(golden["cases"], views):
        assert c["iterations"] == 1_000_000 and c["threads"] == 16 and c["policy"] == "root"
        assert str(modal) in c["best_counts"] and best_by_win in win
        state = capi.KwState.from_buffer_copy(bytes.fromhex(c["state"]))
        legal = capi.kw_legal_mask(state)
        assert all(legal >> int(m) & 1 for m in c["moves"])                 # the reference only visits legal moves
        assert abs(sum(x["visit_share_mean"] for x in c["moves"].values()) - 1.0) < 1e-6


def test_agreement_summary_on_synthetic_results():
    golden = {"cases": [
        {"name": "decisive", "best_counts": {"3": 16}, "reps": 16, "moves": {
            "3": {"visit_share_mean": 0.7, "visit_share_sd": 0.02, "win_rate_mean": 0.40, "win_rate_sd": 0.001},
            "5": {"visit_share_mean": 0.3, "visit_share_sd": 0.02, "win_rate_mean": 0.35, "win_rate_sd": 0.001}}},
        {"name": "open", "best_counts": {"1": 9, "2": 7}, "reps": 16, "moves": {
            "1": {"visit_share_mean": 0.51, "visit_share_sd": 0.05, "win_rate_mean": 0.300, "win_rate_sd": 0.001},
            "2": {"visit_share_mean": 0.49, "visit_share_sd": 0.05, "win_rate_mean": 0.302, "win_rate_sd": 0.001}}}]}
    assert A.reference_view(golden["cases"][0])[:3] == (True, 3, 3)
    assert A.reference_view(golden["cases"][1])[:3] == (False, 1, 2)
    results = [{"best": [3, 3, 3, 5]}, {"best": [1, 1, 2, 2]}]
    s = A.agreement(golden, results)
    assert s["decisive"] == 1 and s["modal_agreement_on_decisive"] == 1.0
    assert s["per_seed_agreement_on_decisive"] == 0.75
    assert abs(s["regret_max"] - 0.05) < 1e-12                              # the one search that played move 5
    assert s["rows"][1]["in_reference_set"] and abs(s["rows"][1]["regret_max"] - 0.002) < 1e-12


def test_position_stream_is_reproducible_and_playable():
    bench = _load_bench()
    a = bench.position_stream(capi, 40)
    b = bench.position_stream(capi, 40)
    assert [bytes(x) for x in a] == [bytes(x) for x in b]
    assert all(capi.kw_legal_mask(s) for s in a)                            # no finished games
    alive = [bin(s.alive_mask).count("1") for s in a]
    assert min(alive) >= 2 and max(alive) == 4 and len({s.tricks_left for s in a}) >= 3   # openings and late rounds


def test_lanes_per_tree_never_below_128_iterations_per_lane():
    # the rule bench.py and cuda_execution.h share: lanes <= iterations / trees / 128
    for trees in (8, 16, 24, 32, 64, 128, 256):
        cap = 1_000_000 // trees // 128
        lanes = max(1, min(128, cap // 32 * 32 if cap >= 32 else cap))
        assert 1_000_000 // trees // lanes >= 128
        assert lanes == 128 or trees > 32


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_harness")), reason="reference harness not built")
def test_reference_seat_plays_legal_moves():
    from ismcsolver_b200.selfplay import ReferenceSeat
    seat = ReferenceSeat(os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "count", 3000, 2)
    try:
        for stream in range(3):
            s = capi.kw_deal(7, stream, 4)
            move = seat.decide(s)
            assert capi.kw_legal_mask(s) >> move & 1
        assert seat.seconds > 0
    finally:
        seat.close()
