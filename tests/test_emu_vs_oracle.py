"""Device logic on the CPU: the kernels' per-thread search body (ismcsolver_b200/csrc/*.cuh), compiled
as plain C++ by tests/emu, against the oracle — so that the bit masks, the hash node pool and the
micro-step game loop are checked here, where there is no GPU.  The real kernels are compared with the
oracle by the `-m gpu` tests; this build is test infrastructure and not part of the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
FIELDS = ("parent", "move", "player", "visits", "score2", "avail")


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["make", "-C", EMU_DIR], check=True, capture_output=True)
    E = C.CDLL(os.path.join(EMU_DIR, "libemu.so"))
    E.emu_playout.restype = C.c_uint32
    E.emu_playout.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
    E.emu_shared_search.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64,
                                    C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    E.emu_search.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32,
                             C.c_uint32, C.c_uint32, C.c_uint32, C.c_double, C.c_uint32, C.c_void_p, C.c_uint32,
                             C.POINTER(C.c_uint32)]
    return E


def emu_search(E, game, root, iters, seed, pos, tree, exploration=0.7, width=1, dump=0):
    """SO-ISMCTS + UCB1; a warp of `width` lanes searches trees tree..tree+width-1, lane `dump` is returned."""
    nodes = np.zeros(iters + 2, dtype=O.NODE_DTYPE)
    n = C.c_uint32()
    assert E.emu_search(game, C.byref(root), O.SOLVER_SO, O.POLICY_UCB1, iters, seed, pos, tree, width, dump,
                        exploration, 0, nodes.ctypes.data, iters + 2, C.byref(n)) == 0
    return nodes[:n.value]


@pytest.mark.parametrize("players", [2, 4, 7])
def test_kw_playouts(emu, players):
    for pos in range(12):
        s = O.kw_deal(1234, pos, players)
        for it in range(100):
            assert O.playout(O.GAME_KW, s, 99, pos, 0, it) == emu.emu_playout(O.GAME_KW, C.byref(s), 99, pos, it)


def test_kw_playouts_midgame(emu):
    rng = np.random.default_rng(0)
    for pos in range(30):
        s = O.kw_deal(8, pos, 4)
        ctr = [0]
        for _ in range(int(rng.integers(0, 60))):
            legal = O.kw_legal(s)
            if not legal:
                break
            O.kw_apply(s, int(rng.choice(legal)), 8, pos, ctr)
        if not O.kw_legal(s):
            continue
        for it in range(50):
            assert O.playout(O.GAME_KW, s, 5, pos, 0, it) == emu.emu_playout(O.GAME_KW, C.byref(s), 5, pos, it)


@pytest.mark.parametrize("mnk", [(3, 3, 3), (4, 7, 4), (15, 15, 5), (16, 16, 5)])
def test_mnk_playouts(emu, mnk):
    s = O.mnk_init(*mnk)
    for it in range(150):
        assert O.playout(O.GAME_MNK, s, 7, 0, 0, it) == emu.emu_playout(O.GAME_MNK, C.byref(s), 7, 0, it)


@pytest.mark.parametrize("players,iters", [(2, 3000), (4, 3000), (7, 1500)])
def test_kw_search_every_node(emu, players, iters):
    for pos in range(3):
        s = O.kw_deal(77, pos, players)
        want = O.so_search(O.GAME_KW, s, iters, seed=5, pos=pos, tree=pos + 1)
        got = emu_search(emu, O.GAME_KW, s, iters, 5, pos, pos + 1)
        assert len(got) == len(want)
        for f in FIELDS:
            assert (got[f] == want[f]).all(), f


@pytest.mark.parametrize("mnk,iters", [((3, 3, 3), 5000), ((7, 7, 4), 3000), ((15, 15, 5), 1500)])
def test_mnk_search_every_node(emu, mnk, iters):
    s = O.mnk_init(*mnk)
    want = O.so_search(O.GAME_MNK, s, iters, seed=5)
    got = emu_search(emu, O.GAME_MNK, s, iters, 5, 0, 0)
    assert len(got) == len(want)
    for f in FIELDS:
        assert (got[f] == want[f]).all(), f


def test_exploration_constant_is_used(emu):
    s = O.kw_deal(1, 1, 4)
    for c in (0.0, 0.35, 2.0):
        want = O.so_search(O.GAME_KW, s, 1500, seed=2, exploration=c)
        got = emu_search(emu, O.GAME_KW, s, 1500, 2, 0, 0, exploration=c)
        for f in FIELDS:
            assert (got[f] == want[f]).all(), (c, f)


@pytest.mark.parametrize("width,dump", [(2, 1), (7, 0), (32, 31), (32, 13)])
def test_a_tree_does_not_depend_on_the_other_lanes_of_its_warp(emu, width, dump):
    """The lock-step loop of lane_search.cuh makes the lanes of a warp wait for each other in every
    phase; what a lane computes must be what it computes alone."""
    s = O.kw_deal(3, 3, 4)
    want = O.so_search(O.GAME_KW, s, 400, seed=8, pos=0, tree=dump)
    got = emu_search(emu, O.GAME_KW, s, 400, 8, 0, 0, width=width, dump=dump)
    assert len(got) == len(want)
    for f in FIELDS:
        assert (got[f] == want[f]).all(), f


def test_a_warp_of_lanes_mid_game_and_seven_players(emu):
    """Lanes whose rounds end at different times (different numbers of players still in)."""
    s = O.kw_deal(11, 2, 7)
    rnd = np.random.default_rng(4)
    ctr = [0]
    for _ in range(37):
        legal = O.kw_legal(s)
        if not legal:
            break
        O.kw_apply(s, int(rnd.choice(legal)), 11, 2, ctr)
    assert O.kw_legal(s)
    for dump in (0, 9):
        want = O.so_search(O.GAME_KW, s, 300, seed=5, pos=0, tree=dump)
        got = emu_search(emu, O.GAME_KW, s, 300, 5, 0, 0, width=12, dump=dump)
        assert len(got) == len(want)
        for f in FIELDS:
            assert (got[f] == want[f]).all(), f


def emu_any(E, game, root, solver, policy, iters, seed, player, width=1, dump=0):
    cap = 8 * (iters + 2)
    nodes = np.zeros(cap, dtype=O.NODE_DTYPE)
    n = C.c_uint32()
    assert E.emu_search(game, C.byref(root), solver, policy, iters, seed, 0, 0, width, dump, 0.7, player,
                        nodes.ctypes.data, cap, C.byref(n)) == 0
    return nodes[:n.value]


EXP_FIELDS = ("parent", "move", "player", "visits", "score", "prob")


@pytest.mark.parametrize("players", [2, 4, 6])
@pytest.mark.parametrize("policy", [O.POLICY_UCB1, O.POLICY_EXP3])
def test_kw_mo_and_exp3_every_node(emu, players, policy):
    """MO-ISMCTS (one tree per player) and EXP3 (as the sequential policy, BASELINE config 4)."""
    s = O.kw_deal(1234, players, players)
    fields = EXP_FIELDS if policy == O.POLICY_EXP3 else FIELDS
    want = O.so_search(O.GAME_KW, s, 1200, seed=5, policy=policy)
    got = emu_any(emu, O.GAME_KW, s, O.SOLVER_SO, policy, 1200, 5, 0)
    assert len(got) == len(want)
    for f in fields:
        assert (got[f] == want[f]).all(), ("so", f)
    trees = O.mo_search(O.GAME_KW, s, 700, seed=6, policy=policy)
    assert sorted(trees) == list(range(players))
    for p, want in trees.items():
        got = emu_any(emu, O.GAME_KW, s, O.SOLVER_MO, policy, 700, 6, p)
        assert len(got) == len(want)
        for f in fields:
            assert (got[f] == want[f]).all(), ("mo", p, f)


@pytest.mark.parametrize("policy", [O.POLICY_UCB1, O.POLICY_EXP3])
def test_mnk_mo_every_node(emu, policy):
    s = O.mnk_init(4, 4, 3)
    fields = EXP_FIELDS if policy == O.POLICY_EXP3 else FIELDS
    trees = O.mo_search(O.GAME_MNK, s, 1500, seed=3, policy=policy)
    for p, want in trees.items():
        got = emu_any(emu, O.GAME_MNK, s, O.SOLVER_MO, policy, 1500, 3, p)
        assert len(got) == len(want)
        for f in fields:
            assert (got[f] == want[f]).all(), (p, f)


@pytest.mark.parametrize("solver,policy", [(O.SOLVER_MO, O.POLICY_UCB1), (O.SOLVER_SO, O.POLICY_EXP3), (O.SOLVER_MO, O.POLICY_EXP3)])
def test_mo_and_exp3_in_a_warp(emu, solver, policy):
    s = O.kw_deal(3, 5, 4)
    fields = FIELDS if policy == O.POLICY_UCB1 else ("parent", "move", "player", "visits", "score", "prob")
    if solver == O.SOLVER_MO:
        want = O.mo_search(O.GAME_KW, s, 250, seed=6, pos=0, tree=3, policy=policy)[1]
    else:
        want = O.so_search(O.GAME_KW, s, 250, seed=6, pos=0, tree=3, policy=policy)
    got = emu_any(emu, O.GAME_KW, s, solver, policy, 250, 6, 1, width=8, dump=3)
    assert len(got) == len(want)
    for f in fields:
        assert (got[f] == want[f]).all(), f


def emu_shared(E, game, root, iters, width, seed, solver=O.SOLVER_SO, policy=O.POLICY_UCB1, player=0):
    nodes = np.zeros(8 * (iters + width + 2), dtype=O.NODE_DTYPE)
    n, done = C.c_uint32(), C.c_uint64()
    assert E.emu_shared_search(game, C.byref(root), solver, policy, iters, width, seed, player, nodes.ctypes.data,
                               len(nodes), C.byref(n), C.byref(done)) == 0
    return nodes[:n.value], done.value


def check_tree_invariants(nodes, iterations, ucb=True):
    """What any correct tree-parallel search satisfies, whatever the interleaving of its lanes."""
    top = nodes[1:][nodes[1:]["parent"] == 0]
    assert int(top["visits"].sum()) == iterations            # every iteration passes through one root child
    assert len(nodes) <= iterations + 1                      # at most one node per iteration
    kids = np.zeros(len(nodes), dtype=np.int64)
    for r in nodes[1:]:
        kids[r["parent"]] += r["visits"]
    assert (kids[1:] <= nodes[1:]["visits"]).all()          # a node is visited at least as often as its children
    assert (nodes[1:]["visits"] >= 1).all()
    if ucb:
        assert (nodes[1:]["score2"] <= 2 * nodes[1:]["visits"]).all()
    else:
        assert (nodes[1:]["score"] >= 0).all() and ((nodes[1:]["prob"] > 0) & (nodes[1:]["prob"] <= 1)).all()
    keys = set(zip(nodes[1:]["parent"].tolist(), nodes[1:]["move"].tolist()))
    assert len(keys) == len(nodes) - 1                       # no duplicated child


def test_shared_tree_one_lane_is_the_sequential_search(emu):
    """Tree parallelisation with one lane in flight = the reference's sequential semantics."""
    for game, s in ((O.GAME_KW, O.kw_deal(5, 5, 4)), (O.GAME_MNK, O.mnk_init(5, 5, 4))):
        want = O.so_search(game, s, 1500, seed=9)
        got, done = emu_shared(emu, game, s, 1500, 1, 9)
        assert done == 1500 and len(got) == len(want)
        for f in FIELDS:
            assert (got[f] == want[f]).all(), f


@pytest.mark.parametrize("width", [2, 7, 32, 128])
def test_shared_tree_invariants(emu, width):
    for game, s in ((O.GAME_KW, O.kw_deal(5, 6, 4)), (O.GAME_MNK, O.mnk_init(4, 4, 3))):
        nodes, done = emu_shared(emu, game, s, 3000, width, 11)
        assert done == 3000
        check_tree_invariants(nodes, 3000)


@pytest.mark.parametrize("solver,policy", [(O.SOLVER_MO, O.POLICY_UCB1), (O.SOLVER_SO, O.POLICY_EXP3), (O.SOLVER_MO, O.POLICY_EXP3)])
def test_shared_trees_mo_and_exp3_one_lane_is_the_sequential_search(emu, solver, policy):
    """MO-ISMCTS and EXP3 on shared trees: with one lane in flight the trees equal the oracle's."""
    s = O.kw_deal(5, 5, 4)
    fields = FIELDS if policy == O.POLICY_UCB1 else ("parent", "move", "player", "visits", "score", "prob")
    for p in ((0, 2) if solver == O.SOLVER_MO else (0,)):
        if solver == O.SOLVER_MO:
            want = O.mo_search(O.GAME_KW, s, 500, seed=9, policy=policy)[p]
        else:
            want = O.so_search(O.GAME_KW, s, 500, seed=9, policy=policy)
        got, done = emu_shared(emu, O.GAME_KW, s, 500, 1, 9, solver, policy, p)
        assert done == 500 and len(got) == len(want)
        for f in fields:
            assert (got[f] == want[f]).all(), (p, f)


@pytest.mark.parametrize("solver,policy", [(O.SOLVER_MO, O.POLICY_UCB1), (O.SOLVER_SO, O.POLICY_EXP3), (O.SOLVER_MO, O.POLICY_EXP3)])
@pytest.mark.parametrize("width", [5, 32])
def test_shared_trees_mo_and_exp3_invariants(emu, solver, policy, width):
    for game, s in ((O.GAME_KW, O.kw_deal(5, 6, 4)), (O.GAME_MNK, O.mnk_init(4, 4, 3))):
        for p in ((0, 1) if solver == O.SOLVER_MO else (0,)):
            nodes, done = emu_shared(emu, game, s, 1500, width, 11, solver, policy, p)
            assert done == 1500
            check_tree_invariants(nodes, 1500, ucb=policy == O.POLICY_UCB1)


def _goof_positions(count, seed=0):
    rnd = np.random.default_rng(seed)
    out = []
    stream = 0
    while len(out) < count:
        s = O.goof_init(17, stream)
        stream += 1
        for _ in range(int(rnd.integers(0, 40))):
            legal = O.goof_legal(s)
            if not legal:
                break
            O.goof_apply(s, int(rnd.choice(legal)))
        if O.goof_legal(s):
            out.append(s)
    return out


def test_goofspiel_playouts(emu):
    for s in _goof_positions(25):
        for it in range(40):
            assert O.playout(O.GAME_GOOF, s, 9, 0, 0, it) == emu.emu_playout(O.GAME_GOOF, C.byref(s), 9, 0, it)


@pytest.mark.parametrize("policy", [O.POLICY_EXP3, O.POLICY_UCB1])
def test_goofspiel_search_every_node(emu, policy):
    """Simultaneous moves -> EXP3 (solverbase.h:58-64); all three observers, incl. the hidden bid seen from
    player 1 (goofspiel.cpp:76-78) and the environment that knows the order of the prizes."""
    fields = EXP_FIELDS if policy == O.POLICY_EXP3 else FIELDS
    seen = set()
    for s in _goof_positions(12, seed=3):
        seen.add(int(s.player))
        want = O.so_search(O.GAME_GOOF, s, 500, seed=4, policy=policy)
        got = emu_any(emu, O.GAME_GOOF, s, O.SOLVER_SO, policy, 500, 4, 0)
        assert len(got) == len(want)
        for f in fields:
            assert (got[f] == want[f]).all(), ("so", f)
        trees = O.mo_search(O.GAME_GOOF, s, 300, seed=5, policy=policy)
        assert sorted(trees) == [0, 1, 2]
        for p, want in trees.items():
            got = emu_any(emu, O.GAME_GOOF, s, O.SOLVER_MO, policy, 300, 5, p)
            assert len(got) == len(want)
            for f in fields:
                assert (got[f] == want[f]).all(), ("mo", p, f)
    assert seen == {0, 1, 2}


def _phantom_positions(count, seed=0):
    rnd = np.random.default_rng(seed)
    out = []
    dims = [(3, 3, 3), (4, 4, 3), (6, 5, 4), (9, 9, 5)]
    while len(out) < count:
        s = O.phantom_init(*dims[len(out) % len(dims)])
        for _ in range(int(rnd.integers(0, 14))):
            legal = O.phantom_legal(s)
            if not legal:
                break
            O.phantom_apply(s, int(rnd.choice(legal)))
        if O.phantom_legal(s):
            out.append(s)
    return out


def test_phantom_playouts(emu):
    for s in _phantom_positions(24):
        for it in range(30):
            assert O.playout(O.GAME_PHANTOM, s, 9, 0, 0, it) == emu.emu_playout(O.GAME_PHANTOM, C.byref(s), 9, 0, it)


@pytest.mark.parametrize("policy", [O.POLICY_UCB1, O.POLICY_EXP3])
def test_phantom_search_every_node(emu, policy):
    """Partially observable moves: the same tree move can be made by different players in different
    determinisations; rewards are credited by the player stored in the node (ucb1.h:53-56)."""
    fields = EXP_FIELDS if policy == O.POLICY_EXP3 else FIELDS
    for s in _phantom_positions(10, seed=2):
        want = O.so_search(O.GAME_PHANTOM, s, 600, seed=4, policy=policy)
        got = emu_any(emu, O.GAME_PHANTOM, s, O.SOLVER_SO, policy, 600, 4, 0)
        assert len(got) == len(want)
        for f in fields:
            assert (got[f] == want[f]).all(), ("so", f)
        for p, want in O.mo_search(O.GAME_PHANTOM, s, 300, seed=5, policy=policy).items():
            got = emu_any(emu, O.GAME_PHANTOM, s, O.SOLVER_MO, policy, 300, 5, p)
            assert len(got) == len(want)
            for f in fields:
                assert (got[f] == want[f]).all(), ("mo", p, f)
