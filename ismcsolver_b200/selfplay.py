"""Batched game driver: many Knockout Whist games advanced in lock-step, every decision of every game
searched in ONE engine call per kind of searcher (the position-batch form of the reference's benchmark
harness, test/benchmark.cpp:73-121: solver against random players, against itself or against another solver,
win statistics and time per move).  Harness code for bench.py, the tools and the tests; the searches are the
engine's."""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class SeatSearch:
    """How a seat decides: `trees` trees per decision sharing `iterations` iterations (count mode) or running
    for `time_ns` each (time mode).  `lanes` > 0: every tree is searched by that many lanes in flight with
    virtual loss (tree parallelisation inside root parallelisation)."""
    trees: int = 1
    iterations: int = 1000
    time_ns: int = 0
    lanes: int = 0
    solver: int = capi.SOLVER_SO
    policy: int = capi.POLICY_UCB1
    exploration: float = 0.7
    max_batch: int = 4096       # positions per engine call (bounds the node pool)
    label: str = ""

    def flags(self) -> int:
        return (capi.FLAG_SHARED_TREE | (self.lanes << capi.FLAG_LANES_SHIFT)) if self.lanes else 0


class ReferenceSeat:
    """A seat played by the UNMODIFIED reference: SOSolver<Card, RootParallel> through oracle/_ref/ref_harness
    (`kw-moves`), one position per request.  Harness code for the strength comparisons only."""

    def __init__(self, harness: str, mode: str, budget: float, threads: int, label: str = "reference"):
        import subprocess
        self.label, self.mode, self.budget, self.threads = label, mode, budget, threads
        self.proc = subprocess.Popen([harness, "kw-moves", mode, str(budget), str(threads)], stdin=subprocess.PIPE,
                                     stdout=subprocess.PIPE, text=True, bufsize=1)
        self.seconds = 0.0

    def decide(self, state) -> int:
        self.proc.stdin.write(bytes(state).hex() + "\n")
        self.proc.stdin.flush()
        move, secs = self.proc.stdout.readline().split()
        self.seconds += float(secs)
        return int(move)

    def close(self):
        try:
            self.proc.stdin.close()
            self.proc.wait(timeout=10)
        except Exception:
            self.proc.kill()


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
    forced: int = 0
    per_seat: dict = field(default_factory=dict)   # seat -> {"decisions", "iterations", "search_seconds"}
    extra: dict = field(default_factory=dict)

    def as_dict(self) -> dict:
        d = {"games": self.games, "players": self.players, "wins": self.wins, "decisions": self.decisions,
             "search_calls": self.search_calls, "iterations": self.iterations, "search_seconds": self.search_seconds,
             "plies": self.plies, "forced_moves_not_searched": self.forced,
             "ms_per_decision": 1e3 * self.search_seconds / max(1, self.decisions),
             "iterations_per_second": self.iterations / max(1e-12, self.search_seconds),
             "per_seat": {str(k): v for k, v in self.per_seat.items()}}
        d.update(self.extra)
        return d


def play_match(engine: capi.Engine, games: int, players: int, seats: dict, seed: int = 1,
               skip_forced: bool = False, first_game: int = 0, on_round=None) -> SelfPlayReport:
    """Plays `games` games (deals `first_game`.. of the deal stream of `seed`).  Seats in `seats` move by search,
    the other seats uniformly at random (benchmark.cpp:78-81).  Every round, all games whose player to move is
    the same searching seat are searched together, one position each.  `skip_forced`: a decision with one
    legal move is played without a search (the move is the same; only the time differs)."""
    rng = np.random.default_rng(seed)
    ids = list(range(first_game, first_game + games))
    states = {g: capi.kw_deal(seed, g, players) for g in ids}
    counters = {g: [0] for g in ids}
    live = list(ids)
    report = SelfPlayReport(games=games, players=players, wins=[0] * players)
    for seat in seats:
        report.per_seat[seat] = {"decisions": 0, "iterations": 0, "search_seconds": 0.0}
    step = 0
    while live:
        # random seats (and forced moves) move until every live game waits for a searching seat (or ends)
        waiting: dict[int, list[int]] = {seat: [] for seat in seats}
        for g in list(live):
            while True:
                mask = capi.kw_legal_mask(states[g])
                if mask == 0:
                    live.remove(g)
                    report.wins[int(np.log2(states[g].alive_mask))] += 1
                    break
                seat = states[g].player
                forced = skip_forced and mask & (mask - 1) == 0
                if seat in seats and not forced:
                    waiting[seat].append(g)
                    break
                legal = [c for c in range(52) if mask >> c & 1]
                capi.kw_apply(states[g], int(rng.choice(legal)), seed, g, counters[g])
                report.plies += 1
                report.forced += int(seat in seats)
        for seat, queue in waiting.items():
            cfg = seats[seat]
            if isinstance(cfg, ReferenceSeat):
                t0 = time.perf_counter()
                for g in queue:
                    move = cfg.decide(states[g])
                    assert capi.kw_legal_mask(states[g]) >> move & 1, "the reference returned an illegal move"
                    capi.kw_apply(states[g], move, seed, g, counters[g])
                    report.plies += 1
                ps = report.per_seat[seat]
                ps["decisions"] += len(queue); ps["search_seconds"] += time.perf_counter() - t0
                continue
            for lo in range(0, len(queue), cfg.max_batch):
                batch = queue[lo:lo + cfg.max_batch]
                desc = engine.make_desc(capi.GAME_KW, len(batch), cfg.trees, iterations=cfg.iterations,
                                        time_ns=cfg.time_ns, seed=seed * 1000003 + step, solver=cfg.solver,
                                        policy=cfg.policy, exploration=cfg.exploration, flags=cfg.flags())
                t0 = time.perf_counter()
                res = engine.search(desc, capi.pack_states([states[g] for g in batch]))
                dt = time.perf_counter() - t0
                done = int(res["iterations_done"].sum())
                report.search_seconds += dt
                report.search_calls += 1
                report.decisions += len(batch)
                report.iterations += done
                ps = report.per_seat[seat]
                ps["decisions"] += len(batch); ps["iterations"] += done; ps["search_seconds"] += dt
                for i, g in enumerate(batch):
                    move = int(res["best_move"][i])
                    assert capi.kw_legal_mask(states[g]) >> move & 1, "the engine returned an illegal move"
                    capi.kw_apply(states[g], move, seed, g, counters[g])
                    report.plies += 1
                step += 1
        if on_round:
            on_round(report, len(live))
    return report


def play_games(engine: capi.Engine, games: int, players: int = 4, solver_seats=(0,), iterations: int = 1000,
               time_ns: int = 0, trees: int = 1, seed: int = 1, solver: int = capi.SOLVER_SO,
               policy: int = capi.POLICY_UCB1, max_batch: int = 4096) -> SelfPlayReport:
    """All searching seats decide the same way (see play_match)."""
    cfg = SeatSearch(trees=trees, iterations=iterations, time_ns=time_ns, solver=solver, policy=policy,
                     max_batch=max_batch)
    return play_match(engine, games, players, {seat: cfg for seat in solver_seats}, seed=seed)
