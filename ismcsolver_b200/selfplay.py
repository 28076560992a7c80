"""Batched game driver: many Knockout Whist games advanced in lock-step, every decision of every game
searched in ONE engine call (the position-batch form of the reference's benchmark harness,
test/benchmark.cpp:73-121: solver against random players or against itself, win statistics and time per
move).  Harness code for bench.py and the tests; the searches are the engine's."""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class SelfPlayReport:
    games: int
    players: int
    wins: list[int]
    decisions: int = 0
    search_calls: int = 0
    iterations: int = 0
    search_seconds: float = 0.0
    plies: int = 0
    extra: dict = field(default_factory=dict)

    def as_dict(self) -> dict:
        d = {"games": self.games, "players": self.players, "wins": self.wins, "decisions": self.decisions,
             "search_calls": self.search_calls, "iterations": self.iterations, "search_seconds": self.search_seconds,
             "plies": self.plies,
             "ms_per_decision": 1e3 * self.search_seconds / max(1, self.decisions),
             "iterations_per_second": self.iterations / max(1e-12, self.search_seconds)}
        d.update(self.extra)
        return d


def play_games(engine: capi.Engine, games: int, players: int = 4, solver_seats=(0,), iterations: int = 1000,
               time_ns: int = 0, trees: int = 1, seed: int = 1, solver: int = capi.SOLVER_SO,
               policy: int = capi.POLICY_UCB1, max_batch: int = 4096) -> SelfPlayReport:
    """Plays `games` games.  Seats in `solver_seats` move by search (iterations per decision, or time_ns
    per decision in time mode), the other seats uniformly at random (benchmark.cpp:78-81).  All games
    whose player to move is a solver seat are searched together, one position each."""
    rng = np.random.default_rng(seed)
    states = [capi.kw_deal(seed, g, players) for g in range(games)]
    counters = [[0] for _ in range(games)]
    live = list(range(games))
    report = SelfPlayReport(games=games, players=players, wins=[0] * players)
    step = 0
    while live:
        # random seats move until every live game waits for a solver seat (or ends)
        waiting = []
        for g in list(live):
            while True:
                mask = capi.kw_legal_mask(states[g])
                if mask == 0:
                    live.remove(g)
                    report.wins[int(np.log2(states[g].alive_mask))] += 1
                    break
                if states[g].player in solver_seats:
                    waiting.append(g)
                    break
                legal = [c for c in range(52) if mask >> c & 1]
                capi.kw_apply(states[g], int(rng.choice(legal)), seed, g, counters[g])
                report.plies += 1
        for lo in range(0, len(waiting), max_batch):
            batch = waiting[lo:lo + max_batch]
            desc = engine.make_desc(capi.GAME_KW, len(batch), trees, iterations=iterations, time_ns=time_ns,
                                    seed=seed * 1000003 + step, solver=solver, policy=policy)
            t0 = time.perf_counter()
            res = engine.search(desc, capi.pack_states([states[g] for g in batch]))
            report.search_seconds += time.perf_counter() - t0
            report.search_calls += 1
            report.decisions += len(batch)
            report.iterations += int(res["iterations_done"].sum())
            for i, g in enumerate(batch):
                move = int(res["best_move"][i])
                assert capi.kw_legal_mask(states[g]) >> move & 1, "the engine returned an illegal move"
                capi.kw_apply(states[g], move, seed, g, counters[g])
                report.plies += 1
            step += 1
    return report
