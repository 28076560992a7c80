#!/usr/bin/env python
"""Generates tests/golden/ref_stats.json from the UNMODIFIED reference (oracle/_ref/ref_harness).

Run in the build container, where /root/reference exists (`make -C oracle` builds the harness).
The reference's random streams cannot be reproduced (unseeded std::mt19937, utility.h:62-67), so
what is recorded is the DISTRIBUTION of its root statistics over many repetitions of the same
search: per root move, mean and standard deviation of the visit share and of the win rate
(score / visits), plus how often each move was returned by operator().  tests/test_ref_stats.py
compares the oracle (and through it the CUDA engine, which is bit-exact with the oracle) against
these numbers.

    python oracle/gen_golden.py            # ~2-3 minutes
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402

SEED = 20261019


def kw_position(stream: int, plies: int, players: int = 4) -> O.KwState:
    """Deal `stream` of the synthetic deal stream advanced by `plies` random legal moves."""
    rng = np.random.default_rng(SEED + stream)
    s = O.kw_deal(SEED, stream, players)
    ctr = [0]
    for _ in range(plies):
        legal = O.kw_legal(s)
        if not legal:
            break
        O.kw_apply(s, int(rng.choice(legal)), SEED, stream, ctr)
    return s


def summarise(rows, stride):
    share = np.zeros((len(rows), stride))
    win = np.full((len(rows), stride), np.nan)
    best = {}
    for i, r in enumerate(rows):
        total = sum(v[0] for v in r["root"].values())
        for mv, (visits, score, _third) in r["root"].items():
            share[i, int(mv)] = visits / total
            if visits > 0:
                win[i, int(mv)] = score / visits
        best[str(r["best"])] = best.get(str(r["best"]), 0) + 1
    moves = {}
    for mv in range(stride):
        if share[:, mv].any():
            w = win[:, mv][~np.isnan(win[:, mv])]
            moves[str(mv)] = {
                "visit_share_mean": float(share[:, mv].mean()), "visit_share_sd": float(share[:, mv].std(ddof=1)),
                "win_rate_mean": float(w.mean()), "win_rate_sd": float(w.std(ddof=1)) if len(w) > 1 else 0.0,
                "present": int((share[:, mv] > 0).sum()),
            }
    return {"moves": moves, "best_counts": best}


def main():
    cases = []
    ref_opening = O.ref_harness("kw-opening", 4).strip()
    kw_cases = [("ref-opening-4p", ref_opening)] + [
        (f"deal{stream}-ply{plies}", kw_position(stream, plies).hex())
        for stream, plies in [(0, 0), (1, 0), (2, 9), (3, 17), (4, 30), (5, 44)]]
    for name, hexstate in kw_cases:
        for solver, policy, it, threads, reps in [("so", "seq", 10000, 1, 30), ("so", "root", 10000, 8, 30)]:
            if name != "ref-opening-4p" and policy == "root" and not name.startswith("deal0"):
                continue
            rows = json.loads(O.ref_harness("kw-stats", hexstate, solver, policy, it, threads, reps))
            cases.append({"name": name, "game": "kw", "state": hexstate, "solver": solver, "policy": policy,
                          "iterations": it, "threads": threads, "reps": reps, **summarise(rows, 64)})
            print(name, solver, policy, cases[-1]["best_counts"], flush=True)
    # MO-ISMCTS (UCB1) and MO with EXP3 as the sequential policy (BASELINE config 4)
    for solver in ("mo", "mo-exp3"):
        rows = json.loads(O.ref_harness("kw-stats", ref_opening, solver, "seq", 5000, 1, 30))
        cases.append({"name": "ref-opening-4p", "game": "kw", "state": ref_opening, "solver": solver, "policy": "seq",
                      "iterations": 5000, "threads": 1, "reps": 30, **summarise(rows, 64)})
        print("ref-opening-4p", solver, cases[-1]["best_counts"], flush=True)
    # m,n,k: empty 3x3 and 5x5 (k=4) boards, and the P1DrawOrLose fixture (solvertest.cpp:45-60)
    import ctypes as C
    for (m, n, k), it in [((3, 3, 3), 5000), ((5, 5, 4), 5000)]:
        s = O.mnk_init(m, n, k)
        rows = json.loads(O.ref_harness("mnk-stats", s.hex(), "so", "seq", it, 1, 30))
        cases.append({"name": f"mnk{m}x{n}x{k}-empty", "game": "mnk", "state": s.hex(), "solver": "so", "policy": "seq",
                      "iterations": it, "threads": 1, "reps": 30, **summarise(rows, 256)})
        print(cases[-1]["name"], cases[-1]["best_counts"], flush=True)
    s = O.mnk_init(3, 3, 3)
    for mv in (4, 1, 5, 3, 6, 8, 7):
        O.lib().orc_mnk_apply(C.byref(s), mv)
    for solver in ("so", "mo"):
        rows = json.loads(O.ref_harness("mnk-stats", s.hex(), solver, "seq", 16, 1, 100))
        cases.append({"name": "p1-draw-or-lose", "game": "mnk", "state": s.hex(), "solver": solver, "policy": "seq",
                      "iterations": 16, "threads": 1, "reps": 100, **summarise(rows, 256)})
        print("p1-draw-or-lose", solver, cases[-1]["best_counts"], flush=True)
    # Goofspiel: every move is simultaneous -> EXP3 (solverbase.h:58-64).  Player 0 to move after the first
    # prize, and player 1 to move after player 0's (hidden) bid (goofspiel.cpp:76-78).
    for prize, bid, solver in [(11, -1, "so"), (11, 10, "so"), (5, -1, "mo"), (0, 3, "mo")]:
        g = O.GoofState()
        rest = [r for r in range(13) if r != prize]
        g.prize_order = sum(r << (4 * i) for i, r in enumerate(rest))
        g.n_prizes = 12
        g.hands[0] = g.hands[1] = 0x1FFF
        g.current_prize = prize
        g.player, g.draw_prize = 0, 0
        if bid >= 0:
            g.moves[0] = bid
            g.player = 1
        rows = json.loads(O.ref_harness("goof-stats", prize, bid, solver, 3000, 60))
        cases.append({"name": f"goof-prize{prize}-bid{bid}", "game": "goof", "state": g.hex(), "solver": solver + "-sim",
                      "policy": "seq", "iterations": 3000, "threads": 1, "reps": 60, **summarise(rows, 64)})
        print(cases[-1]["name"], solver, cases[-1]["best_counts"], flush=True)
    # phantom m,n,k: positions reached by move sequences; the hidden stones are too few to complete a line, so
    # the reference's "ensure a non-winning state" replay (phantommnkgame.cpp:57-81) never restarts and its
    # determinisation is exactly the uniform one
    for name, dims, seq in [("phantom3-gametest", (3, 3, 3), [4, 4, 0, 6, 2]),        # gametest.cpp:264-279
                            ("phantom4", (4, 4, 3), [5, 5, 10, 6, 0, 9]),
                            ("phantom5", (5, 5, 4), [12, 12, 6, 7, 18, 8, 13, 16])]:
        s = O.phantom_init(*dims)
        for mv in seq:
            assert O.phantom_apply(s, mv) == 0
        for solver, it in (("so", 3000), ("mo", 3000)):
            rows = json.loads(O.ref_harness("phantom-stats", s.hex(), solver, "seq", it, 1, 30))
            cases.append({"name": name, "game": "phantom", "state": s.hex(), "solver": solver, "policy": "seq",
                          "iterations": it, "threads": 1, "reps": 30, **summarise(rows, 256)})
            print(name, solver, cases[-1]["best_counts"], flush=True)
    # the reference's own game code: uniform random playouts from its opening position
    play = json.loads(O.ref_harness("kw-playouts", ref_opening, 200000))
    out = {"generator": "oracle/gen_golden.py", "reference": "sfranzen/ismcsolver (unmodified, g++ -O3 -DNDEBUG)",
           "cases": cases, "kw_playouts": {"state": ref_opening, **play}}
    path = os.path.join(ROOT, "tests", "golden", "ref_stats.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
