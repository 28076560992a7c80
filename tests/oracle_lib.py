"""ctypes binding of the CPU oracle (oracle/liboracle.so) — test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
module; nothing under ismcsolver_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_HARNESS = os.path.join(ORACLE_DIR, "_ref", "ref_harness")

GAME_KW, GAME_MNK, GAME_GOOF, GAME_PHANTOM = 0, 1, 2, 3
SOLVER_SO, SOLVER_MO = 0, 1
POLICY_UCB1, POLICY_EXP3 = 0, 1


class KwState(C.Structure):
    """Mirror of ismcts_kw_state (include/ismcts_b200.h)."""
    _fields_ = [
        ("hands", C.c_uint64 * 7),
        ("unknown", C.c_uint64),
        ("num_players", C.c_uint8),
        ("player", C.c_uint8),
        ("dealer", C.c_uint8),
        ("trump", C.c_uint8),
        ("tricks_left", C.c_uint8),
        ("request_trump", C.c_uint8),
        ("alive_mask", C.c_uint8),
        ("trick_len", C.c_uint8),
        ("tricks_taken", C.c_uint8 * 8),
        ("trick_player", C.c_uint8 * 8),
        ("trick_card", C.c_uint8 * 8),
    ]

    def copy(self) -> "KwState":
        out = KwState()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(KwState))
        return out

    def hex(self) -> str:
        return bytes(self).hex()

    @staticmethod
    def from_hex(text: str) -> "KwState":
        return KwState.from_buffer_copy(bytes.fromhex(text))


class MnkState(C.Structure):
    """Mirror of ismcts_mnk_state."""
    _fields_ = [
        ("stones", (C.c_uint64 * 4) * 2),
        ("m", C.c_uint8),
        ("n", C.c_uint8),
        ("k", C.c_uint8),
        ("player", C.c_uint8),
        ("result2", C.c_int8),
        ("pad", C.c_uint8 * 3),
    ]

    def copy(self) -> "MnkState":
        out = MnkState()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(MnkState))
        return out

    def hex(self) -> str:
        return bytes(self).hex()


class GoofState(C.Structure):
    """Mirror of ismcts_goof_state."""
    _fields_ = [
        ("prize_order", C.c_uint64),
        ("hands", C.c_uint16 * 2),
        ("n_prizes", C.c_uint8),
        ("player", C.c_uint8),
        ("draw_prize", C.c_uint8),
        ("current_prize", C.c_uint8),
        ("moves", C.c_uint8 * 2),
        ("scores", C.c_uint8 * 2),
        ("pad", C.c_uint8 * 4),
    ]

    def copy(self) -> "GoofState":
        out = GoofState()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(GoofState))
        return out

    def hex(self) -> str:
        return bytes(self).hex()


class PhantomState(C.Structure):
    """Mirror of ismcts_phantom_state."""
    _fields_ = [
        ("stones", (C.c_uint64 * 4) * 2),
        ("avail", (C.c_uint64 * 4) * 2),
        ("m", C.c_uint8), ("n", C.c_uint8), ("k", C.c_uint8), ("player", C.c_uint8),
        ("result2", C.c_int8), ("pad", C.c_uint8 * 3),
    ]

    def copy(self) -> "PhantomState":
        out = PhantomState()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(PhantomState))
        return out

    def hex(self) -> str:
        return bytes(self).hex()


assert C.sizeof(KwState) == 96 and C.sizeof(MnkState) == 72 and C.sizeof(GoofState) == 24
assert C.sizeof(PhantomState) == 136

NODE_DTYPE = np.dtype([
    ("parent", "<u4"), ("move", "<u4"), ("player", "<u4"), ("visits", "<u4"),
    ("score2", "<u4"), ("avail", "<u4"), ("score", "<f8"), ("prob", "<f8"),
])
assert NODE_DTYPE.itemsize == 40


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "iterations", "plies", "draws", "chance_picks", "select_steps", "child_evals",
        "nodes_created", "backprop_updates")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build_oracle(force: bool = False) -> str:
    """Compiles oracle/liboracle.so (and oracle/_ref when the reference checkout exists)."""
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("oracle.c", "oracle.h")] + [
        os.path.join(ROOT, "include", f) for f in ("ismcts_b200.h", "ismcts_detmath.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
    return so


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        u32, u64, dbl, vp = C.c_uint32, C.c_uint64, C.c_double, C.c_void_p
        L.orc_draw_word.restype = u32
        L.orc_draw_word.argtypes = [u64, u32, u32, u32, u32]
        L.orc_kw_deal.argtypes = [u64, u32, u32, C.POINTER(KwState)]
        L.orc_kw_legal.restype = u64
        L.orc_kw_legal.argtypes = [C.POINTER(KwState)]
        L.orc_kw_apply.argtypes = [C.POINTER(KwState), u32, u64, u32, C.POINTER(u32)]
        L.orc_kw_result2.argtypes = [C.POINTER(KwState), u32]
        L.orc_kw_determinise.argtypes = [C.POINTER(KwState), u32, u64, u32, u32, u32, C.POINTER(u32)]
        L.orc_mnk_init.argtypes = [u32, u32, u32, C.POINTER(MnkState)]
        L.orc_mnk_legal.argtypes = [C.POINTER(MnkState), C.POINTER(u64)]
        L.orc_mnk_apply.argtypes = [C.POINTER(MnkState), u32]
        L.orc_mnk_result2.argtypes = [C.POINTER(MnkState), u32]
        L.orc_phantom_init.argtypes = [u32, u32, u32, C.POINTER(PhantomState)]
        L.orc_phantom_legal.argtypes = [C.POINTER(PhantomState), C.POINTER(u64)]
        L.orc_phantom_legal.restype = None
        L.orc_phantom_apply.argtypes = [C.POINTER(PhantomState), u32]
        L.orc_phantom_result2.argtypes = [C.POINTER(PhantomState), u32]
        L.orc_goof_init.argtypes = [u64, u32, C.POINTER(GoofState)]
        L.orc_goof_legal.restype = u64
        L.orc_goof_legal.argtypes = [C.POINTER(GoofState)]
        L.orc_goof_apply.argtypes = [C.POINTER(GoofState), u32]
        L.orc_goof_result2.argtypes = [C.POINTER(GoofState), u32]
        L.orc_playout.restype = u32
        L.orc_playout.argtypes = [u32, vp, u64, u32, u32, u32]
        L.orc_so_search.argtypes = [u32, vp, u32, dbl, u64, u32, u32, u64, vp, u32, C.POINTER(u32),
                                    C.POINTER(Counters)]
        L.orc_mo_search.argtypes = [u32, vp, u32, dbl, u64, u32, u32, u64, vp, u32, C.POINTER(u32),
                                    C.POINTER(u32), C.POINTER(Counters)]
        L.orc_root_parallel.argtypes = [u32, u32, vp, u32, dbl, u64, u32, u32, u32, u32, u64, vp, vp, vp,
                                        C.POINTER(u32), C.POINTER(Counters)]
        L.orc_tree_iterations.restype = u64
        L.orc_tree_iterations.argtypes = [u64, u32, u32]
        L.orc_ucb.restype = dbl
        L.orc_ucb.argtypes = [u32, u32, u32, dbl]
        L.orc_exp3_probabilities.argtypes = [vp, vp, u32, vp]
        _lib = L
    return _lib


def move_stride(game: int) -> int:
    return 256 if game in (GAME_MNK, GAME_PHANTOM) else 64


def phantom_init(m: int = 3, n: int = 3, k: int = 3) -> PhantomState:
    s = PhantomState()
    assert lib().orc_phantom_init(m, n, k, C.byref(s)) == 0
    return s


def phantom_legal(s: PhantomState) -> list[int]:
    legal = (C.c_uint64 * 4)()
    lib().orc_phantom_legal(C.byref(s), legal)
    return [c for c in range(256) if legal[c >> 6] >> (c & 63) & 1]


def phantom_apply(s: PhantomState, move: int) -> int:
    return lib().orc_phantom_apply(C.byref(s), move)


def phantom_result2(s: PhantomState, player: int) -> int:
    return lib().orc_phantom_result2(C.byref(s), player)


def goof_init(seed: int, stream: int = 0) -> GoofState:
    s = GoofState()
    lib().orc_goof_init(seed, stream, C.byref(s))
    return s


def goof_legal(s: GoofState) -> list[int]:
    mask = lib().orc_goof_legal(C.byref(s))
    return [c for c in range(64) if mask >> c & 1]


def goof_apply(s: GoofState, move: int) -> int:
    return lib().orc_goof_apply(C.byref(s), move)


def goof_result2(s: GoofState, player: int) -> int:
    return lib().orc_goof_result2(C.byref(s), player)


def kw_deal(seed: int, stream: int, players: int = 4) -> KwState:
    s = KwState()
    lib().orc_kw_deal(seed, stream, players, C.byref(s))
    return s


def kw_legal(s: KwState) -> list[int]:
    mask = lib().orc_kw_legal(C.byref(s))
    return [c for c in range(52) if mask >> c & 1]


def kw_apply(s: KwState, move: int, seed: int = 0, stream: int = 0, counter: list | None = None) -> int:
    d = C.c_uint32(counter[0] if counter else 0)
    rc = lib().orc_kw_apply(C.byref(s), move, seed, stream, C.byref(d))
    if counter is not None:
        counter[0] = d.value
    return rc


def mnk_init(m: int = 3, n: int = 3, k: int = 3) -> MnkState:
    s = MnkState()
    assert lib().orc_mnk_init(m, n, k, C.byref(s)) == 0
    return s


def mnk_legal(s: MnkState) -> list[int]:
    legal = (C.c_uint64 * 4)()
    lib().orc_mnk_legal(C.byref(s), legal)
    return [c for c in range(256) if legal[c >> 6] >> (c & 63) & 1]


def playout(game: int, root, seed: int, pos: int, tree: int, iteration: int) -> int:
    return lib().orc_playout(game, C.byref(root), seed, pos, tree, iteration)


def so_search(game, root, iterations, seed=1, pos=0, tree=0, policy=POLICY_UCB1, exploration=0.7,
              counters: Counters | None = None) -> np.ndarray:
    """One sequential tree; returns the node records in creation order."""
    cap = int(iterations) + 2
    nodes = np.zeros(cap, dtype=NODE_DTYPE)
    n = C.c_uint32(0)
    rc = lib().orc_so_search(game, C.byref(root), policy, exploration, seed, pos, tree, iterations,
                             nodes.ctypes.data, cap, C.byref(n), C.byref(counters) if counters else None)
    assert rc == 0, rc
    return nodes[: n.value].copy()


def mo_search(game, root, iterations, seed=1, pos=0, tree=0, policy=POLICY_UCB1, exploration=0.7,
              counters: Counters | None = None) -> dict[int, np.ndarray]:
    """MO-ISMCTS: returns {player: node records of that player's tree}."""
    cap = 8 * (int(iterations) + 2)
    nodes = np.zeros(cap, dtype=NODE_DTYPE)
    offs = (C.c_uint32 * 8)()
    cnts = (C.c_uint32 * 8)()
    rc = lib().orc_mo_search(game, C.byref(root), policy, exploration, seed, pos, tree, iterations,
                             nodes.ctypes.data, cap, offs, cnts, C.byref(counters) if counters else None)
    assert rc == 0, rc
    return {p: nodes[offs[p]: offs[p] + cnts[p]].copy() for p in range(8) if cnts[p]}


def root_parallel(game, root, iterations, trees, seed=1, pos=0, tree_base=0, trees_total=None,
                  solver=SOLVER_SO, policy=POLICY_UCB1, exploration=0.7, counters: Counters | None = None):
    """Returns (visits, score2, avail, best_move) summed over the given trees."""
    stride = move_stride(game)
    v = np.zeros(stride, dtype=np.uint64)
    s = np.zeros(stride, dtype=np.uint64)
    a = np.zeros(stride, dtype=np.uint64)
    best = C.c_uint32(0)
    rc = lib().orc_root_parallel(game, solver, C.byref(root), policy, exploration, seed, pos, tree_base, trees,
                                 trees_total or (tree_base + trees), iterations, v.ctypes.data, s.ctypes.data,
                                 a.ctypes.data, C.byref(best), C.byref(counters) if counters else None)
    assert rc == 0, rc
    return v, s, a, best.value


def ref_harness(*args: str, timeout: float = 600.0) -> str:
    """Runs oracle/_ref/ref_harness (the unmodified reference) and returns its stdout."""
    out = subprocess.run([REF_HARNESS, *map(str, args)], check=True, capture_output=True, text=True,
                         timeout=timeout)
    return out.stdout
