"""Engine decisions against the golden root statistics of the unmodified reference (tests/golden/ref_stats_1m.json:
SOSolver<Card, RootParallel>{1M, 16}, 16 repetitions per position).  Harness code for tests/test_gpu_geometry.py and
tools/geometry_agreement.py; the searches are the engine's.

A position is DECISIVE when the gap between the visit shares of the reference's two most visited root moves exceeds
three standard deviations of that gap over its repetitions.  The REGRET of a decision is measured with
the reference's own estimator: the reference's mean root win rate of its best move minus that of the move decided on.
"""
from __future__ import annotations

import json
import math
import os
from collections import Counter

import numpy as np

from . import capi

GOLDEN_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_stats_1m.json")


def load_golden():
    return json.load(open(GOLDEN_PATH))


def search_golden_positions(engine: capi.Engine, golden, trees: int, lanes: int = 0, seeds: int = 8, seed: int = 1234):
    """Every golden position `seeds` times (different position ids = different Philox streams) in one engine call.
    Returns per position: visit shares [seeds, 64], win rates [seeds, 64], decisions [seeds]."""
    cases = golden["cases"]
    states = [capi.KwState.from_buffer_copy(bytes.fromhex(c["state"])) for c in cases for _ in range(seeds)]
    flags = (capi.FLAG_SHARED_TREE | (lanes << capi.FLAG_LANES_SHIFT)) if lanes else 0
    desc = engine.make_desc(capi.GAME_KW, len(states), trees, iterations=cases[0]["iterations"], seed=seed, flags=flags)
    res = engine.search(desc, capi.pack_states(states))
    assert int(res["iterations_done"].sum()) == len(states) * cases[0]["iterations"]
    out = []
    for i in range(len(cases)):
        rows = slice(i * seeds, (i + 1) * seeds)
        v = res["visits"][rows].astype(np.float64)
        s = res["score2"][rows].astype(np.float64) * 0.5
        out.append({"share": v / v.sum(axis=1, keepdims=True),
                    "win": np.divide(s, v, out=np.zeros_like(s), where=v > 0),
                    "best": [int(b) for b in res["best_move"][rows]]})
    return out


def reference_view(case):
    """(decisive?, the reference's most frequent decision, its best move by win rate, win rate per move)"""
    moves = {int(m): x for m, x in case["moves"].items()}
    by_share = sorted(moves, key=lambda m: -moves[m]["visit_share_mean"])
    decisive = True
    if len(by_share) > 1:
        a, b = moves[by_share[0]], moves[by_share[1]]
        gap = a["visit_share_mean"] - b["visit_share_mean"]
        decisive = gap > 3.0 * math.hypot(a["visit_share_sd"], b["visit_share_sd"])
    modal = int(max(case["best_counts"], key=lambda m: case["best_counts"][m]))
    win = {m: x["win_rate_mean"] for m, x in moves.items()}
    return decisive, modal, max(win, key=win.get), win


def agreement(golden, results) -> dict:
    """Summary of `results` (search_golden_positions) against the golden cases."""
    rows = []
    for case, got in zip(golden["cases"], results):
        decisive, modal, best_by_win, win = reference_view(case)
        mine = Counter(got["best"]).most_common(1)[0][0]
        regrets = [win[best_by_win] - win.get(b, 0.0) for b in got["best"]]
        rows.append({"name": case["name"], "decisive": decisive, "reference": modal, "engine": mine,
                     "engine_decisions": got["best"], "agree_modal": mine == modal,
                     "agree_per_seed": sum(b == modal for b in got["best"]) / len(got["best"]),
                     "in_reference_set": all(str(b) in case["best_counts"] for b in got["best"]),
                     "regret_max": max(regrets), "regret_mean": float(np.mean(regrets))})
    dec = [r for r in rows if r["decisive"]]
    return {"positions": len(rows), "decisive": len(dec),
            "modal_agreement_on_decisive": sum(r["agree_modal"] for r in dec) / max(1, len(dec)),
            "per_seed_agreement_on_decisive": float(np.mean([r["agree_per_seed"] for r in dec])) if dec else 1.0,
            "modal_agreement_all": sum(r["agree_modal"] for r in rows) / len(rows),
            "regret_max": max(r["regret_max"] for r in rows), "regret_mean": float(np.mean([r["regret_mean"] for r in rows])),
            "rows": rows}
