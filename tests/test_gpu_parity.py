"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle (bit-exact).

The oracle (oracle/oracle.c) restates the reference's algorithm with different data structures
(membership arrays and child lists instead of bit masks and a hash pool) and is driven by the same
Philox streams, so every integer the search produces must agree.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from ismcsolver_b200 import capi

pytestmark = pytest.mark.gpu

FIELDS = ("parent", "move", "player", "visits", "score2", "avail")
import json as _json
import os as _os
GOLDEN_OPENING = _json.load(open(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden",
                                               "ref_stats.json")))["cases"][0]["state"]


def _okw(s):
    return O.KwState.from_buffer_copy(bytes(s))


def _omnk(s):
    return O.MnkState.from_buffer_copy(bytes(s))


def _midgame(seed, stream, plies):
    """Position stream: `plies` uniformly random legal moves from deal `stream`."""
    rng = np.random.default_rng(seed * 1000003 + stream)
    s = capi.kw_deal(seed, stream, 4)
    ctr = [0]
    for _ in range(plies):
        mask = capi.kw_legal_mask(s)
        legal = [c for c in range(52) if mask >> c & 1]
        if not legal:
            break
        assert capi.kw_apply(s, int(rng.choice(legal)), seed, stream, ctr) == 0
    return s


def test_playouts_kw_bit_exact(engine):
    """Determinise + rollout outcomes (winner, plies, draws) for 20k (position, iteration) pairs."""
    n_pos, per = 40, 512
    states = [_midgame(11, i, (i * 7) % 50) for i in range(n_pos)]
    states = [s for s in states if capi.kw_legal_mask(s)]
    got = engine.playouts(capi.GAME_KW, capi.pack_states(states), len(states), per, seed=4242)
    for i, s in enumerate(states):
        os_ = _okw(s)
        want = np.array([O.playout(O.GAME_KW, os_, 4242, i, 0, j) for j in range(per)], dtype=np.uint32)
        assert (got[i] == want).all(), f"position {i}"
    # every playout ends with exactly one survivor (getResult sums to 1)
    assert ((got & 0xFF) < 4).all()


@pytest.mark.parametrize("mnk", [(3, 3, 3), (5, 4, 3), (9, 9, 5), (15, 15, 5)])
def test_playouts_mnk_bit_exact(engine, mnk):
    s = capi.mnk_init(*mnk)
    got = engine.playouts(capi.GAME_MNK, capi.pack_states([s]), 1, 1024, seed=9)
    want = np.array([O.playout(O.GAME_MNK, _omnk(s), 9, 0, 0, j) for j in range(1024)], dtype=np.uint32)
    assert (got[0] == want).all()


def test_so_tree_kw_every_node_bit_exact(engine):
    """visits / score / available / parent / move / player of EVERY node of several trees."""
    n_pos, trees, per_tree = 3, 4, 1500
    states = [_midgame(5, i, i * 9) for i in range(n_pos)]
    desc = engine.make_desc(capi.GAME_KW, n_pos, trees, iterations=trees * per_tree, seed=77)
    res = engine.search(desc, capi.pack_states(states))
    assert (res["iterations_done"] == per_tree).all()
    for i, s in enumerate(states):
        for t in range(trees):
            got = engine.dump_tree(i * trees + t, per_tree + 2)
            want = O.so_search(O.GAME_KW, _okw(s), per_tree, seed=77, pos=i, tree=t)
            assert len(got) == len(want)
            for f in FIELDS:
                assert (got[f] == want[f]).all(), (i, t, f)


@pytest.mark.parametrize("mnk,iters", [((3, 3, 3), 4000), ((6, 5, 4), 2000), ((15, 15, 5), 1000)])
def test_so_tree_mnk_every_node_bit_exact(engine, mnk, iters):
    s = capi.mnk_init(*mnk)
    desc = engine.make_desc(capi.GAME_MNK, 1, 2, iterations=2 * iters, seed=3)
    engine.search(desc, capi.pack_states([s]))
    for t in range(2):
        got = engine.dump_tree(t, iters + 2)
        want = O.so_search(O.GAME_MNK, _omnk(s), iters, seed=3, pos=0, tree=t)
        assert len(got) == len(want)
        for f in FIELDS:
            assert (got[f] == want[f]).all(), (t, f)


def test_root_parallel_sums_and_best_move(engine):
    """RootParallel::compileVisitCounts + bestMove (execution.h:208-231) over 64 trees, uneven split."""
    n_pos, trees, iterations = 6, 64, 64 * 100 + 17
    states = [_midgame(21, i, i * 5) for i in range(n_pos)]
    desc = engine.make_desc(capi.GAME_KW, n_pos, trees, iterations=iterations, seed=123, pos_base=1000)
    res = engine.search(desc, capi.pack_states(states))
    assert int(res["iterations_done"].sum()) == n_pos * iterations
    for i, s in enumerate(states):
        v, sc, av, best = O.root_parallel(O.GAME_KW, _okw(s), iterations, trees, seed=123, pos=1000 + i)
        assert (res["visits"][i] == v).all()
        assert (res["score2"][i] == sc).all()
        assert (res["avail"][i] == av).all()
        assert int(res["best_move"][i]) == best
        assert int(res["visits"][i].sum()) == iterations  # every iteration passes through one root child
        assert capi.kw_legal_mask(s) >> best & 1           # solvertest.cpp:109-111: a valid move


def test_tree_sharding_is_invariant(engine):
    """Splitting the tree range over several calls (as over several GPUs) gives identical sums."""
    s = capi.kw_deal(8, 0, 4)
    roots = capi.pack_states([s])
    total_trees, iterations = 32, 32 * 64 + 5
    whole = engine.search(engine.make_desc(capi.GAME_KW, 1, total_trees, iterations=iterations, seed=9), roots)
    acc = {k: np.zeros_like(whole[k]) for k in ("visits", "score2", "avail")}
    for base in range(0, total_trees, 8):
        part = engine.search(engine.make_desc(capi.GAME_KW, 1, 8, iterations=iterations, seed=9,
                                              trees_total=total_trees, tree_base=base), roots)
        for k in acc:
            acc[k] += part[k]
    for k in acc:
        assert (acc[k] == whole[k]).all()


def test_full_size_decision_properties(engine):
    """BASELINE.json configs[1] at its full size (1M iterations per decision over 2368 trees): what does not
    depend on the size — every iteration is accounted for, passes through exactly one legal root child,
    scores never exceed visits, two half-launches with the tree range split (as two GPUs would) give the
    sums of the whole launch, and the same launch twice gives the same numbers."""
    n_pos, trees, iterations = 4, 2368, 1_000_000
    states = [capi.kw_deal(0x1535C75B200, i, 4) for i in range(n_pos)]
    roots = capi.pack_states(states)
    whole = engine.search(engine.make_desc(capi.GAME_KW, n_pos, trees, iterations=iterations, seed=0xB200), roots)
    assert int(whole["iterations_done"].sum()) == n_pos * iterations
    per_tree = whole["iterations_done"]
    assert int(per_tree.max()) - int(per_tree.min()) <= 1          # execution.h:42-52: an even split
    for i, s in enumerate(states):
        legal = capi.kw_legal_mask(s)
        assert int(whole["visits"][i].sum()) == iterations
        for m in range(64):
            if not legal >> m & 1:
                assert whole["visits"][i][m] == 0
        assert (whole["score2"][i] <= 2 * whole["visits"][i]).all()
        assert legal >> int(whole["best_move"][i]) & 1
    halves = {k: np.zeros_like(whole[k]) for k in ("visits", "score2", "avail")}
    for base in (0, trees // 2):
        part = engine.search(engine.make_desc(capi.GAME_KW, n_pos, trees // 2, iterations=iterations, seed=0xB200,
                                              trees_total=trees, tree_base=base), roots)
        for k in halves:
            halves[k] += part[k]
    again = engine.search(engine.make_desc(capi.GAME_KW, n_pos, trees, iterations=iterations, seed=0xB200), roots)
    for k in halves:
        assert (again[k] == whole[k]).all() and (halves[k] == whole[k]).all(), k
    # one tree of the full-size launch against the oracle, every node
    it0 = int(per_tree[0, 5])
    nodes = O.so_search(O.GAME_KW, _okw(states[0]), it0, seed=0xB200, pos=0, tree=5)
    got = engine.dump_tree(5, it0 + 2)
    assert len(got) == len(nodes)
    for f in FIELDS:
        assert (got[f] == nodes[f]).all(), f


def test_p1_draw_or_lose(engine):
    """solvertest.cpp:45-60,131-138: from P1DrawOrLose every solver must play move 2 with 16 iterations."""
    s = capi.mnk_init(3, 3, 3)
    for mv in (4, 1, 5, 3, 6, 8, 7):  # X: 4,5,6,7  O: 1,3,8 -> board [[-,O,-],[O,X,X],[X,X,O]]
        assert capi.mnk_apply(s, mv) == 0
    assert capi.mnk_legal(s) == [0, 2] and s.player == 1
    for seed in range(20):
        res = engine.search(engine.make_desc(capi.GAME_MNK, 1, 1, iterations=16, seed=seed), capi.pack_states([s]))
        assert int(res["best_move"][0]) == 2


def test_terminal_root_and_errors(engine):
    # a finished game has no legal move: best_move is the sentinel and no iteration adds a node
    s = capi.mnk_init(3, 3, 3)
    for mv in (0, 3, 1, 4, 2):
        assert capi.mnk_apply(s, mv) == 0
    assert capi.mnk_legal(s) == []
    res = engine.search(engine.make_desc(capi.GAME_MNK, 1, 2, iterations=10, seed=1), capi.pack_states([s]))
    assert int(res["best_move"][0]) == 0xFFFFFFFF and int(res["visits"].sum()) == 0
    with pytest.raises(capi.EngineError):
        engine.search(engine.make_desc(7, 1, 1, iterations=10), capi.pack_states([s]))
    with pytest.raises(capi.EngineError):
        engine.search(engine.make_desc(capi.GAME_MNK, 1, 0, iterations=10), capi.pack_states([s]))


def test_time_mode_runs_and_replays(engine):
    """Time-limited search (execution.h:188, utility.h:51-60): iterations done are reported per tree
    and the oracle replays exactly those counts."""
    s = capi.kw_deal(3, 1, 4)
    trees = 8
    res = engine.search(engine.make_desc(capi.GAME_KW, 1, trees, time_ns=5_000_000, seed=55), capi.pack_states([s]))
    done = res["iterations_done"][0]
    assert (done > 0).all()
    v = np.zeros(64, dtype=np.uint64)
    for t in range(trees):
        nodes = O.so_search(O.GAME_KW, _okw(s), int(done[t]), seed=55, pos=0, tree=t)
        for nrec in nodes[1:]:
            if nrec["parent"] == 0:
                v[nrec["move"]] += nrec["visits"]
    assert (v == res["visits"][0]).all()


EXP_FIELDS = ("parent", "move", "player", "visits", "score", "prob")


@pytest.mark.parametrize("policy", [capi.POLICY_UCB1, capi.POLICY_EXP3])
def test_mo_trees_kw_every_node_bit_exact(engine, policy):
    """MO-ISMCTS (mosolver.h:57-101): every player's tree of several lanes; EXP3 as the sequential
    policy is BASELINE config 4 (MOSolver<Card, Policy, EXP3>)."""
    n_pos, lanes, per_lane = 2, 3, 600
    states = [_midgame(9, i, i * 11) for i in range(n_pos)]
    desc = engine.make_desc(capi.GAME_KW, n_pos, lanes, iterations=lanes * per_lane, seed=31, solver=capi.SOLVER_MO,
                            policy=policy)
    res = engine.search(desc, capi.pack_states(states))
    assert (res["iterations_done"] == per_lane).all()
    fields = EXP_FIELDS if policy == capi.POLICY_EXP3 else FIELDS
    for i, s in enumerate(states):
        for t in range(lanes):
            want = O.mo_search(O.GAME_KW, _okw(s), per_lane, seed=31, pos=i, tree=t, policy=policy)
            for player, w in want.items():
                got = engine.dump_tree(i * lanes + t, 4 * per_lane, player=player)
                assert len(got) == len(w)
                for f in fields:
                    assert (got[f] == w[f]).all(), (i, t, player, f)
        # the decision comes from the tree of the root's player to move (mosolver.h:42-46)
        v = np.zeros(64, dtype=np.uint64)
        for t in range(lanes):
            w = O.mo_search(O.GAME_KW, _okw(s), per_lane, seed=31, pos=i, tree=t, policy=policy)[s.player]
            for nrec in w[1:]:
                if nrec["parent"] == 0:
                    v[nrec["move"]] += nrec["visits"]
        assert (v == res["visits"][i]).all()
        assert int(res["best_move"][i]) == int(np.argmax(v))


def test_so_exp3_and_mnk_mo_bit_exact(engine):
    s = capi.kw_deal(4, 4, 4)
    per = 900
    engine.search(engine.make_desc(capi.GAME_KW, 1, 2, iterations=2 * per, seed=2, policy=capi.POLICY_EXP3),
                  capi.pack_states([s]))
    for t in range(2):
        got = engine.dump_tree(t, per + 2)
        want = O.so_search(O.GAME_KW, _okw(s), per, seed=2, tree=t, policy=O.POLICY_EXP3)
        assert len(got) == len(want)
        for f in EXP_FIELDS:
            assert (got[f] == want[f]).all(), (t, f)
    m = capi.mnk_init(5, 5, 4)
    engine.search(engine.make_desc(capi.GAME_MNK, 1, 1, iterations=1200, seed=8, solver=capi.SOLVER_MO),
                  capi.pack_states([m]))
    want = O.mo_search(O.GAME_MNK, _omnk(m), 1200, seed=8)
    for player, w in want.items():
        got = engine.dump_tree(0, 3000, player=player)
        assert len(got) == len(w)
        for f in FIELDS:
            assert (got[f] == w[f]).all(), (player, f)


# ------------------------------------------------------------------ tree parallelisation

SHARED = 1  # ISMCTS_FLAG_SHARED_TREE


def _shared_flags(lanes):
    return SHARED | (lanes << 8)


def _check_tree_invariants(nodes, iterations):
    top = nodes[1:][nodes[1:]["parent"] == 0]
    assert int(top["visits"].sum()) == iterations
    assert len(nodes) <= iterations + 1
    kids = np.zeros(len(nodes), dtype=np.int64)
    for r in nodes[1:]:
        kids[r["parent"]] += r["visits"]
    assert (kids[1:] <= nodes[1:]["visits"]).all()
    assert (nodes[1:]["score2"] <= 2 * nodes[1:]["visits"]).all() and (nodes[1:]["visits"] >= 1).all()
    assert len(set(zip(nodes[1:]["parent"].tolist(), nodes[1:]["move"].tolist()))) == len(nodes) - 1


def test_shared_tree_one_lane_is_bit_exact(engine):
    """CudaTreeParallel with one iteration in flight is the sequential search (every node)."""
    for game, s, conv in ((capi.GAME_KW, capi.kw_deal(5, 5, 4), _okw), (capi.GAME_MNK, capi.mnk_init(6, 6, 4), _omnk)):
        desc = engine.make_desc(game, 1, 1, iterations=2000, seed=9, flags=_shared_flags(1))
        res = engine.search(desc, capi.pack_states([s]))
        assert int(res["iterations_done"][0, 0]) == 2000
        got = engine.dump_tree(0, 2100)
        want = O.so_search(game, conv(s), 2000, seed=9)
        assert len(got) == len(want)
        for f in FIELDS:
            assert (got[f] == want[f]).all(), f


@pytest.mark.parametrize("lanes", [32, 128, 1024, 8192])
def test_shared_tree_invariants_and_decision(engine, lanes):
    """Many lanes on one tree with virtual loss (atomics): the tree is consistent whatever the
    interleaving, and on a decisive position the decision is the sequential one."""
    iterations = 60000
    s = O.KwState.from_hex(GOLDEN_OPENING)      # the reference's own opening deal: K of hearts (37) in 40/40 runs
    desc = engine.make_desc(capi.GAME_KW, 1, 1, iterations=iterations, seed=4, flags=_shared_flags(lanes))
    res = engine.search(desc, np.frombuffer(bytes(s), dtype=np.uint8))
    assert int(res["iterations_done"].sum()) == iterations
    nodes = engine.dump_tree(0, iterations + 2)
    _check_tree_invariants(nodes, iterations)
    assert int(res["visits"][0].sum()) == iterations
    assert int(res["best_move"][0]) == 37


def test_shared_tree_mnk_p1_draw_or_lose(engine):
    s = capi.mnk_init(3, 3, 3)
    for mv in (4, 1, 5, 3, 6, 8, 7):
        assert capi.mnk_apply(s, mv) == 0
    for seed in range(10):
        res = engine.search(engine.make_desc(capi.GAME_MNK, 1, 1, iterations=16, seed=seed, flags=_shared_flags(4)),
                            capi.pack_states([s]))
        assert int(res["best_move"][0]) == 2
    # larger budget, many lanes: still the drawing move
    res = engine.search(engine.make_desc(capi.GAME_MNK, 1, 1, iterations=20000, seed=1, flags=_shared_flags(2048)),
                        capi.pack_states([s]))
    assert int(res["best_move"][0]) == 2 and int(res["visits"][0].sum()) == 20000


def test_shared_tree_rejects_zero_lanes(engine):
    s = capi.kw_deal(1, 1, 4)
    with pytest.raises(capi.EngineError):
        engine.search(engine.make_desc(capi.GAME_KW, 1, 1, iterations=100, flags=SHARED), capi.pack_states([s]))


@pytest.mark.parametrize("solver,policy", [(capi.SOLVER_MO, capi.POLICY_UCB1), (capi.SOLVER_SO, capi.POLICY_EXP3),
                                           (capi.SOLVER_MO, capi.POLICY_EXP3)])
def test_shared_trees_mo_and_exp3_one_lane_is_bit_exact(engine, solver, policy):
    """MO-ISMCTS (the trees of all players shared) and EXP3 on shared trees: one lane in flight is the
    sequential search, every node."""
    s = capi.kw_deal(5, 5, 4)
    fields = FIELDS if policy == capi.POLICY_UCB1 else ("parent", "move", "player", "visits", "score", "prob")
    desc = engine.make_desc(capi.GAME_KW, 1, 1, iterations=800, seed=9, solver=solver, policy=policy,
                            flags=_shared_flags(1))
    res = engine.search(desc, capi.pack_states([s]))
    assert int(res["iterations_done"][0, 0]) == 800
    for p in ((0, 1, 3) if solver == capi.SOLVER_MO else (0,)):
        got = engine.dump_tree(0, 1000, player=p)
        if solver == capi.SOLVER_MO:
            want = O.mo_search(O.GAME_KW, _okw(s), 800, seed=9, policy=policy)[p]
        else:
            want = O.so_search(O.GAME_KW, _okw(s), 800, seed=9, policy=policy)
        assert len(got) == len(want)
        for f in fields:
            assert (got[f] == want[f]).all(), (p, f)


@pytest.mark.parametrize("solver,policy", [(capi.SOLVER_MO, capi.POLICY_UCB1), (capi.SOLVER_SO, capi.POLICY_EXP3),
                                           (capi.SOLVER_MO, capi.POLICY_EXP3)])
@pytest.mark.parametrize("lanes", [64, 4096])
def test_shared_trees_mo_and_exp3_many_lanes(engine, solver, policy, lanes):
    """Many lanes on the shared trees: every tree stays consistent, every iteration is accounted for and the
    decision on the reference's decisive opening deal is the sequential one."""
    iterations = 40000
    s = O.KwState.from_hex(GOLDEN_OPENING)
    desc = engine.make_desc(capi.GAME_KW, 1, 1, iterations=iterations, seed=4, solver=solver, policy=policy,
                            flags=_shared_flags(lanes))
    res = engine.search(desc, np.frombuffer(bytes(s), dtype=np.uint8))
    assert int(res["iterations_done"].sum()) == iterations
    assert int(res["visits"][0].sum()) == iterations
    for p in ((0, 2) if solver == capi.SOLVER_MO else (0,)):
        nodes = engine.dump_tree(0, 4 * iterations + 8, player=p)
        top = nodes[1:][nodes[1:]["parent"] == 0]
        assert int(top["visits"].sum()) == iterations
        kids = np.zeros(len(nodes), dtype=np.int64)
        for r in nodes[1:]:
            kids[r["parent"]] += r["visits"]
        assert (kids[1:] <= nodes[1:]["visits"]).all() and (nodes[1:]["visits"] >= 1).all()
        assert len(set(zip(nodes[1:]["parent"].tolist(), nodes[1:]["move"].tolist()))) == len(nodes) - 1
    if policy == capi.POLICY_UCB1:
        assert int(res["best_move"][0]) == 37


# ------------------------------------------------------------------ Goofspiel (simultaneous moves, EXP3)

def _goof_positions(count, seed=0):
    rnd = np.random.default_rng(seed)
    out, stream = [], 0
    while len(out) < count:
        s = capi.goof_init(17, stream)
        stream += 1
        for _ in range(int(rnd.integers(0, 40))):
            legal = capi.goof_legal(s)
            if not legal:
                break
            capi.goof_apply(s, int(rnd.choice(legal)))
        if capi.goof_legal(s):
            out.append(s)
    return out


def _ogoof(s):
    return O.GoofState.from_buffer_copy(bytes(s))


def test_goofspiel_playouts_and_trees_bit_exact(engine):
    states = _goof_positions(9, seed=3)
    assert {int(s.player) for s in states} == {0, 1, 2}
    got = engine.playouts(capi.GAME_GOOF, capi.pack_states(states), len(states), 256, seed=11)
    for i, s in enumerate(states):
        want = np.array([O.playout(O.GAME_GOOF, _ogoof(s), 11, i, 0, j) for j in range(256)], dtype=np.uint32)
        assert (got[i] == want).all(), i
    per = 700
    for solver in (capi.SOLVER_SO, capi.SOLVER_MO):
        desc = engine.make_desc(capi.GAME_GOOF, len(states), 2, iterations=2 * per, seed=6, solver=solver,
                                policy=capi.POLICY_EXP3)
        res = engine.search(desc, capi.pack_states(states))
        assert (res["iterations_done"] == per).all()
        for i, s in enumerate(states):
            for t in range(2):
                if solver == capi.SOLVER_SO:
                    wants = {None: O.so_search(O.GAME_GOOF, _ogoof(s), per, seed=6, pos=i, tree=t, policy=O.POLICY_EXP3)}
                else:
                    wants = O.mo_search(O.GAME_GOOF, _ogoof(s), per, seed=6, pos=i, tree=t, policy=O.POLICY_EXP3)
                for player, want in wants.items():
                    got_nodes = engine.dump_tree(i * 2 + t, 4 * per, player=player)
                    assert len(got_nodes) == len(want)
                    for f in EXP_FIELDS:
                        assert (got_nodes[f] == want[f]).all(), (solver, i, t, player, f)
            assert int(res["best_move"][i]) in capi.goof_legal(s)        # solvertest.cpp:121-128


# ------------------------------------------------------------------ phantom m,n,k (partially observable moves)

def test_phantom_playouts_and_trees_bit_exact(engine):
    rnd = np.random.default_rng(4)
    states = []
    dims = [(3, 3, 3), (4, 4, 3), (6, 5, 4), (9, 9, 5)]
    while len(states) < 8:
        s = capi.phantom_init(*dims[len(states) % 4])
        for _ in range(int(rnd.integers(0, 14))):
            legal = capi.phantom_legal(s)
            if not legal:
                break
            capi.phantom_apply(s, int(rnd.choice(legal)))
        if capi.phantom_legal(s):
            states.append(s)
    conv = lambda s: O.PhantomState.from_buffer_copy(bytes(s))
    got = engine.playouts(capi.GAME_PHANTOM, capi.pack_states(states), len(states), 128, seed=13)
    for i, s in enumerate(states):
        want = np.array([O.playout(O.GAME_PHANTOM, conv(s), 13, i, 0, j) for j in range(128)], dtype=np.uint32)
        assert (got[i] == want).all(), i
    per = 500
    for solver in (capi.SOLVER_SO, capi.SOLVER_MO):
        desc = engine.make_desc(capi.GAME_PHANTOM, len(states), 2, iterations=2 * per, seed=8, solver=solver)
        res = engine.search(desc, capi.pack_states(states))
        assert (res["iterations_done"] == per).all()
        for i, s in enumerate(states):
            for t in range(2):
                if solver == capi.SOLVER_SO:
                    wants = {None: O.so_search(O.GAME_PHANTOM, conv(s), per, seed=8, pos=i, tree=t)}
                else:
                    wants = O.mo_search(O.GAME_PHANTOM, conv(s), per, seed=8, pos=i, tree=t)
                for player, want in wants.items():
                    nodes = engine.dump_tree(i * 2 + t, 4 * per + 8, player=player)
                    assert len(nodes) == len(want)
                    for f in FIELDS:
                        assert (nodes[f] == want[f]).all(), (solver, i, t, player, f)
            assert int(res["best_move"][i]) in capi.phantom_legal(s)


def test_persistent_grid_recycles_tables(engine):
    """More private trees than the GPU holds lanes: the warps of a persistent grid claim trees and reuse their
    tables (engine.cu: search_kernel).  The root sums must equal the oracle's and those of the same search with
    one table per tree (ISMCTS_FLAG_KEEP_TREES); afterwards the trees cannot be dumped."""
    n_pos, trees = 5, 40000                       # 200,000 trees > 148 SMs x 8 blocks x 128 lanes
    iterations = trees * 5 + 3
    states = [_midgame(33, i, i * 6) for i in range(n_pos)]
    roots = capi.pack_states(states)
    recycled = engine.search(engine.make_desc(capi.GAME_KW, n_pos, trees, iterations=iterations, seed=404), roots)
    with pytest.raises(capi.EngineError):
        engine.dump_tree(0, 16)
    kept = engine.search(engine.make_desc(capi.GAME_KW, n_pos, trees, iterations=iterations, seed=404,
                                          flags=capi.FLAG_KEEP_TREES), roots)
    assert len(engine.dump_tree(n_pos * trees - 1, 16)) >= 2
    for k in ("visits", "score2", "avail", "iterations_done", "best_move"):
        assert (recycled[k] == kept[k]).all(), k
    assert int(recycled["iterations_done"].sum()) == n_pos * iterations
    for i in (0, n_pos - 1):
        v, sc, av, best = O.root_parallel(O.GAME_KW, _okw(states[i]), iterations, trees, seed=404, pos=i)
        assert (recycled["visits"][i] == v).all() and (recycled["score2"][i] == sc).all()
        assert (recycled["avail"][i] == av).all() and int(recycled["best_move"][i]) == best
    # MO-ISMCTS + EXP3 tables (96-byte slots, seven roots) through the same path
    mo = [engine.search(engine.make_desc(capi.GAME_KW, 2, 90000, iterations=90000 * 3, seed=11, solver=capi.SOLVER_MO,
                                         policy=capi.POLICY_EXP3, flags=f), roots[:2 * 96])
          for f in (0, capi.FLAG_KEEP_TREES)]
    for k in ("visits", "score2", "avail", "iterations_done"):
        assert (mo[0][k] == mo[1][k]).all(), k


def test_tree_parallel_lanes_cap_keeps_the_decision(engine):
    """Tree parallelisation at 1M iterations on the empty 9x9 board (k = 5): with the lanes capped at
    iterations / 128 (cuda_execution.h: lanesFor) all of eight independent searches play the centre, as the
    sequential search does (profiles/r02_tree_parallel_quality.md); with 151,552 lanes in flight — 6.6 iterations
    per lane, what round 1 benchmarked — the search has lost its focus: the most visited move holds less than half
    the share of the visits it holds at the cap (measured 0.06 against 0.26; sequential 0.40) and some of the
    searches decide otherwise (which ones depends on how the hardware interleaves the lanes)."""
    iterations, seeds, centre = 1_000_000, 8, 40
    roots = capi.pack_states([capi.mnk_init(9, 9, 5)] * seeds)

    def search(lanes):
        res = engine.search(engine.make_desc(capi.GAME_MNK, seeds, 1, iterations=iterations, seed=20262,
                                             flags=_shared_flags(lanes)), roots)
        assert (res["iterations_done"].sum(axis=1) == iterations).all()
        share = (res["visits"].max(axis=1) / res["visits"].sum(axis=1)).mean()
        return [int(b) for b in res["best_move"]], float(share)

    capped, capped_share = search(iterations // 128 // 32 * 32)
    assert capped == [centre] * seeds, capped
    assert capped_share >= 0.15, capped_share
    _, flooded_share = search(151552)
    assert flooded_share <= 0.5 * capped_share, (flooded_share, capped_share)


def test_malformed_root_records_are_rejected(engine):
    """Everything a kernel indexes with is checked on the host before a search or a playout: a record
    with a player, a trick entry, a hand or a board out of range fails with ISMCTS_ERR_INVALID and names the position."""
    good = capi.kw_deal(1, 1, 4)

    def broken(**fields):
        s = capi.KwState.from_buffer_copy(bytes(good))
        for k, v in fields.items():
            setattr(s, k, v)
        return s

    cases = [broken(num_players=8), broken(num_players=1), broken(player=4), broken(dealer=7), broken(alive_mask=0),
             broken(alive_mask=0b10001), broken(trump=4), broken(tricks_left=8), broken(trick_len=5)]
    s = capi.KwState.from_buffer_copy(bytes(good)); s.hands[0] = (1 << 8) - 1; cases.append(s)          # eight cards
    s = capi.KwState.from_buffer_copy(bytes(good)); s.hands[1] |= 1 << 60; cases.append(s)               # a card that does not exist
    s = capi.KwState.from_buffer_copy(bytes(good)); s.hands[5] = 1; cases.append(s)                      # a player that does not exist
    s = capi.KwState.from_buffer_copy(bytes(good)); s.trick_len = 1; s.trick_player[0] = 6; cases.append(s)
    for i, bad in enumerate(cases):
        roots = capi.pack_states([good, bad])
        with pytest.raises(capi.EngineError) as e:
            engine.search(engine.make_desc(capi.GAME_KW, 2, 2, iterations=20, seed=1), roots)
        assert e.value.code == capi.ERR_INVALID and "position 1" in str(e.value), (i, str(e.value))
        with pytest.raises(capi.EngineError):
            engine.playouts(capi.GAME_KW, roots, 2, 4, seed=1)
    board = capi.mnk_init(3, 3, 3)
    board.m, board.n = 20, 20                                                                            # 400 cells
    with pytest.raises(capi.EngineError):
        engine.search(engine.make_desc(capi.GAME_MNK, 1, 1, iterations=10), capi.pack_states([board]))
    # unknown flag bits, and lanes = 0 with a shared tree
    with pytest.raises(capi.EngineError):
        engine.search(engine.make_desc(capi.GAME_KW, 1, 1, iterations=10, flags=16), capi.pack_states([good]))
    with pytest.raises(capi.EngineError):
        engine.search(engine.make_desc(capi.GAME_KW, 1, 1, iterations=10, flags=capi.FLAG_SHARED_TREE), capi.pack_states([good]))
