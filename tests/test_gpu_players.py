"""GPU parity for Knockout Whist with 2, 3, 5, 6 and 7 players (the reference clamps to 2..7,
knockoutwhist.cpp:38): positions with more than four players run `search_kernel<KwGameT<7>, ...>` — seven
seats of scratch per lane, longer paths — which no four-player test reaches.  Every comparison is bit-exact
against the CPU oracle through the C ABI, as in test_gpu_parity.py."""
import numpy as np
import pytest

import oracle_lib as O
from ismcsolver_b200 import capi

pytestmark = pytest.mark.gpu

FIELDS = ("parent", "move", "player", "visits", "score2", "avail")
EXP_FIELDS = ("parent", "move", "player", "visits", "score", "prob")
PLAYERS = [2, 3, 5, 6, 7]


def _okw(s):
    return O.KwState.from_buffer_copy(bytes(s))


def _position(seed, stream, players, plies):
    """`plies` uniformly random legal moves from deal `stream` with `players` players."""
    rng = np.random.default_rng(seed * 1000003 + stream * 101 + players)
    s = capi.kw_deal(seed, stream, players)
    ctr = [0]
    for _ in range(plies):
        mask = capi.kw_legal_mask(s)
        legal = [c for c in range(52) if mask >> c & 1]
        if not legal:
            break
        assert capi.kw_apply(s, int(rng.choice(legal)), seed, stream, ctr) == 0
    return s


def _knocked_out_root(players):
    """A mid-game root of a `players`-player deal in which at least one player has been knocked out (two
    players: a knock-out ends the game, so a root in a late round instead)."""
    if players == 2:
        for stream in range(200):
            for plies in range(20, 56, 3):
                s = _position(41, stream, players, plies)
                if capi.kw_legal_mask(s) and s.tricks_left <= 4:
                    return s
    for stream in range(200):
        for plies in range(players * 7, players * 28, 3):
            s = _position(41, stream, players, plies)
            alive = bin(s.alive_mask).count("1")
            if capi.kw_legal_mask(s) and 2 <= alive < players:
                return s
    raise AssertionError("no such position found")


@pytest.mark.parametrize("players", PLAYERS)
def test_playouts_bit_exact_and_turn_bound(engine, players):
    """Outcomes (winner, plies, draws) of determinise + rollout; a game never takes more than 6 + players * 28
    turns (gametest.cpp:166-178)."""
    per = 256
    states = [_position(13, i, players, (i * 5) % (players * 12)) for i in range(12)] + [_knocked_out_root(players)]
    states = [s for s in states if capi.kw_legal_mask(s)]
    got = engine.playouts(capi.GAME_KW, capi.pack_states(states), len(states), per, seed=99)
    for i, s in enumerate(states):
        want = np.array([O.playout(O.GAME_KW, _okw(s), 99, i, 0, j) for j in range(per)], dtype=np.uint32)
        assert (got[i] == want).all(), f"position {i}"
    assert ((got & 0xFF) < players).all()
    # plies from the opening positions (every third state starts at ply 0 when (i * 5) % ... == 0)
    opening = engine.playouts(capi.GAME_KW, capi.pack_states([capi.kw_deal(13, 500 + i, players) for i in range(4)]),
                              4, 2048, seed=5)
    assert int(((opening >> 8) & 0xFFF).max()) <= 6 + players * 28


@pytest.mark.parametrize("players", PLAYERS)
def test_so_tree_every_node_bit_exact(engine, players):
    n_pos, trees, per_tree = 3, 3, 1200
    states = [_position(7, 0, players, 0), _position(7, 1, players, players * 4), _knocked_out_root(players)]
    desc = engine.make_desc(capi.GAME_KW, n_pos, trees, iterations=trees * per_tree, seed=17)
    res = engine.search(desc, capi.pack_states(states))
    assert (res["iterations_done"] == per_tree).all()
    for i, s in enumerate(states):
        for t in range(trees):
            got = engine.dump_tree(i * trees + t, per_tree + 2)
            want = O.so_search(O.GAME_KW, _okw(s), per_tree, seed=17, pos=i, tree=t)
            assert len(got) == len(want)
            for f in FIELDS:
                assert (got[f] == want[f]).all(), (players, i, t, f)


@pytest.mark.parametrize("players", PLAYERS)
def test_root_parallel_sums_and_best_move(engine, players):
    n_pos, trees, iterations = 4, 48, 48 * 90 + 11
    states = [_position(23, i, players, i * players) for i in range(n_pos - 1)] + [_knocked_out_root(players)]
    desc = engine.make_desc(capi.GAME_KW, n_pos, trees, iterations=iterations, seed=321, pos_base=50)
    res = engine.search(desc, capi.pack_states(states))
    for i, s in enumerate(states):
        v, sc, av, best = O.root_parallel(O.GAME_KW, _okw(s), iterations, trees, seed=321, pos=50 + i)
        assert (res["visits"][i] == v).all() and (res["score2"][i] == sc).all() and (res["avail"][i] == av).all()
        assert int(res["best_move"][i]) == best
        assert capi.kw_legal_mask(s) >> best & 1


@pytest.mark.parametrize("players", [3, 7])
@pytest.mark.parametrize("policy", [capi.POLICY_UCB1, capi.POLICY_EXP3])
def test_mo_trees_every_node_bit_exact(engine, players, policy):
    lanes, per_lane = 2, 400
    s = _position(29, 3, players, players * 3)
    desc = engine.make_desc(capi.GAME_KW, 1, lanes, iterations=lanes * per_lane, seed=37, solver=capi.SOLVER_MO,
                            policy=policy)
    engine.search(desc, capi.pack_states([s]))
    fields = EXP_FIELDS if policy == capi.POLICY_EXP3 else FIELDS
    for t in range(lanes):
        want = O.mo_search(O.GAME_KW, _okw(s), per_lane, seed=37, pos=0, tree=t, policy=policy)
        assert len(want) == bin(s.alive_mask).count("1")
        for player, w in want.items():
            got = engine.dump_tree(t, 8 * per_lane, player=player)
            assert len(got) == len(w)
            for f in fields:
                assert (got[f] == w[f]).all(), (players, t, player, f)


def test_mixed_player_counts_in_one_call(engine):
    """Positions with different numbers of players in one batch (the seven-seat kernel serves them all)."""
    states = [_position(31, p, p, p * 2) for p in (2, 3, 4, 5, 6, 7)]
    trees, per_tree = 2, 500
    res = engine.search(engine.make_desc(capi.GAME_KW, len(states), trees, iterations=trees * per_tree, seed=3),
                        capi.pack_states(states))
    for i, s in enumerate(states):
        v, sc, av, best = O.root_parallel(O.GAME_KW, _okw(s), trees * per_tree, trees, seed=3, pos=i)
        assert (res["visits"][i] == v).all() and (res["score2"][i] == sc).all() and (res["avail"][i] == av).all()


def test_time_mode_on_a_forced_root_stays_inside_the_ln_table(engine):
    """A root with ONE legal move in time mode: the tree stops growing but the availability of its nodes keeps
    counting; with many small tables the iteration count would pass the end of the ln() table.
    The search must stop there, return the forced move, and the oracle must replay the reported counts."""
    # the last trick of the last round between the two players left: player 0 holds one card, so does player 2
    forced = capi.KwState()
    forced.num_players, forced.alive_mask, forced.player, forced.dealer = 4, 0b0101, 0, 2
    forced.trump, forced.tricks_left = 1, 1
    forced.hands[0], forced.hands[2] = 1 << 5, 1 << 20
    forced.unknown = ((1 << 52) - 1) & ~(1 << 5) & ~(1 << 20)
    trees = 148 * 8 * 128
    res = engine.search(engine.make_desc(capi.GAME_KW, 1, trees, time_ns=60_000_000, seed=77), capi.pack_states([forced]))
    move = int(res["best_move"][0])
    assert capi.kw_legal_mask(forced) == 1 << move
    done = res["iterations_done"][0]
    assert int(res["visits"][0][move]) == int(done.sum())
    t = int(np.argmax(done))
    nodes = O.so_search(O.GAME_KW, _okw(forced), int(done[t]), seed=77, pos=0, tree=t)
    got = engine.dump_tree(t, len(nodes) + 2)
    assert len(got) == len(nodes)
    for f in FIELDS:
        assert (got[f] == nodes[f]).all(), f
