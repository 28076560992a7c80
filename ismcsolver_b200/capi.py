"""ctypes view of the C ABI (include/ismcts_b200.h) for the Python harnesses (tests, bench).

The product's host layer is C++ (include/ismcts/*.h, the reference's own language); Python is
only plumbing here.  Nothing in this module computes a search: it fails loudly when the
library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB_PATH

GAME_KW, GAME_MNK, GAME_GOOF, GAME_PHANTOM = 0, 1, 2, 3
SOLVER_SO, SOLVER_MO = 0, 1
POLICY_UCB1, POLICY_EXP3 = 0, 1
OK, ERR_NO_DEVICE, ERR_INVALID, ERR_ILLEGAL_MOVE, ERR_CUDA, ERR_CAPACITY = range(6)
NO_DUMP = 0xFFFFFFFF
FLAG_SHARED_TREE, FLAG_KW_MAX4, FLAG_KEEP_TREES, FLAG_LANES_SHIFT = 1, 2, 4, 8

EXPORTS = [
    "ismcts_state_size", "ismcts_move_stride",
    "ismcts_kw_deal", "ismcts_kw_legal", "ismcts_kw_apply", "ismcts_kw_result2",
    "ismcts_mnk_init", "ismcts_mnk_legal", "ismcts_mnk_apply", "ismcts_mnk_result2",
    "ismcts_goof_init", "ismcts_goof_legal", "ismcts_goof_apply", "ismcts_goof_result2",
    "ismcts_phantom_init", "ismcts_phantom_legal", "ismcts_phantom_apply", "ismcts_phantom_result2",
    "ismcts_create", "ismcts_destroy", "ismcts_last_error", "ismcts_set_stream",
    "ismcts_search", "ismcts_search_device", "ismcts_search_multi", "ismcts_synchronize", "ismcts_dump_tree", "ismcts_dump_tree_of",
    "ismcts_playouts", "ismcts_launch_count", "ismcts_version",
]


class KwState(C.Structure):
    """ismcts_kw_state"""
    _fields_ = [
        ("hands", C.c_uint64 * 7), ("unknown", C.c_uint64),
        ("num_players", C.c_uint8), ("player", C.c_uint8), ("dealer", C.c_uint8), ("trump", C.c_uint8),
        ("tricks_left", C.c_uint8), ("request_trump", C.c_uint8), ("alive_mask", C.c_uint8),
        ("trick_len", C.c_uint8),
        ("tricks_taken", C.c_uint8 * 8), ("trick_player", C.c_uint8 * 8), ("trick_card", C.c_uint8 * 8),
    ]


class MnkState(C.Structure):
    """ismcts_mnk_state"""
    _fields_ = [
        ("stones", (C.c_uint64 * 4) * 2),
        ("m", C.c_uint8), ("n", C.c_uint8), ("k", C.c_uint8), ("player", C.c_uint8),
        ("result2", C.c_int8), ("pad", C.c_uint8 * 3),
    ]


class GoofState(C.Structure):
    """ismcts_goof_state"""
    _fields_ = [
        ("prize_order", C.c_uint64), ("hands", C.c_uint16 * 2), ("n_prizes", C.c_uint8), ("player", C.c_uint8),
        ("draw_prize", C.c_uint8), ("current_prize", C.c_uint8), ("moves", C.c_uint8 * 2), ("scores", C.c_uint8 * 2),
        ("pad", C.c_uint8 * 4),
    ]


class PhantomState(C.Structure):
    """ismcts_phantom_state"""
    _fields_ = [
        ("stones", (C.c_uint64 * 4) * 2), ("avail", (C.c_uint64 * 4) * 2),
        ("m", C.c_uint8), ("n", C.c_uint8), ("k", C.c_uint8), ("player", C.c_uint8),
        ("result2", C.c_int8), ("pad", C.c_uint8 * 3),
    ]


class SearchDesc(C.Structure):
    """ismcts_search_desc"""
    _fields_ = [
        ("game", C.c_uint32), ("solver", C.c_uint32), ("policy", C.c_uint32), ("n_positions", C.c_uint32),
        ("trees", C.c_uint32), ("trees_total", C.c_uint32), ("tree_base", C.c_uint32), ("pos_base", C.c_uint32),
        ("iterations", C.c_uint64), ("time_ns", C.c_uint64), ("seed", C.c_uint64),
        ("exploration", C.c_double), ("dump_tree", C.c_uint32), ("flags", C.c_uint32),
    ]


class SearchOut(C.Structure):
    """ismcts_search_out"""
    _fields_ = [
        ("visits", C.c_void_p), ("score2", C.c_void_p), ("avail", C.c_void_p), ("iterations_done", C.c_void_p),
        ("best_move", C.c_void_p), ("dump_nodes", C.c_void_p), ("dump_capacity", C.c_uint32),
        ("dump_count", C.c_uint32),
    ]


NODE_DTYPE = np.dtype([
    ("parent", "<u4"), ("move", "<u4"), ("player", "<u4"), ("visits", "<u4"),
    ("score2", "<u4"), ("avail", "<u4"), ("score", "<f8"), ("prob", "<f8"),
])
assert C.sizeof(KwState) == 96 and C.sizeof(MnkState) == 72 and NODE_DTYPE.itemsize == 40

_lib = None


def load_library() -> C.CDLL:
    """Loads libismcts_b200.so.  Raises when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the engine has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    u32, u64, vp, i32 = C.c_uint32, C.c_uint64, C.c_void_p, C.c_int
    L.ismcts_state_size.restype = C.c_size_t
    L.ismcts_state_size.argtypes = [u32]
    L.ismcts_move_stride.restype = u32
    L.ismcts_move_stride.argtypes = [u32]
    L.ismcts_kw_deal.argtypes = [u64, u32, u32, C.POINTER(KwState)]
    L.ismcts_kw_legal.restype = u64
    L.ismcts_kw_legal.argtypes = [C.POINTER(KwState)]
    L.ismcts_kw_apply.argtypes = [C.POINTER(KwState), u32, u64, u32, C.POINTER(u32)]
    L.ismcts_kw_result2.argtypes = [C.POINTER(KwState), u32]
    L.ismcts_mnk_init.argtypes = [u32, u32, u32, C.POINTER(MnkState)]
    L.ismcts_mnk_legal.restype = None
    L.ismcts_mnk_legal.argtypes = [C.POINTER(MnkState), C.POINTER(u64)]
    L.ismcts_mnk_apply.argtypes = [C.POINTER(MnkState), u32]
    L.ismcts_mnk_result2.argtypes = [C.POINTER(MnkState), u32]
    L.ismcts_phantom_init.argtypes = [u32, u32, u32, C.POINTER(PhantomState)]
    L.ismcts_phantom_legal.restype = None
    L.ismcts_phantom_legal.argtypes = [C.POINTER(PhantomState), C.POINTER(u64)]
    L.ismcts_phantom_apply.argtypes = [C.POINTER(PhantomState), u32]
    L.ismcts_phantom_result2.argtypes = [C.POINTER(PhantomState), u32]
    L.ismcts_goof_init.argtypes = [u64, u32, C.POINTER(GoofState)]
    L.ismcts_goof_legal.restype = u64
    L.ismcts_goof_legal.argtypes = [C.POINTER(GoofState)]
    L.ismcts_goof_apply.argtypes = [C.POINTER(GoofState), u32]
    L.ismcts_goof_result2.argtypes = [C.POINTER(GoofState), u32]
    L.ismcts_create.argtypes = [i32, C.POINTER(vp)]
    L.ismcts_destroy.restype = None
    L.ismcts_destroy.argtypes = [vp]
    L.ismcts_last_error.restype = C.c_char_p
    L.ismcts_last_error.argtypes = [vp]
    L.ismcts_set_stream.argtypes = [vp, vp]
    L.ismcts_search.argtypes = [vp, C.POINTER(SearchDesc), vp, C.POINTER(SearchOut)]
    L.ismcts_search_device.argtypes = [vp, C.POINTER(SearchDesc), vp, vp, vp, vp, vp]
    L.ismcts_search_multi.argtypes = [C.POINTER(vp), u32, C.POINTER(SearchDesc), vp, C.POINTER(SearchOut)]
    L.ismcts_synchronize.argtypes = [vp]
    L.ismcts_dump_tree.argtypes = [vp, u32, vp, u32, C.POINTER(u32)]
    L.ismcts_dump_tree_of.argtypes = [vp, u32, u32, vp, u32, C.POINTER(u32)]
    L.ismcts_playouts.argtypes = [vp, u32, vp, u32, u32, u64, vp]
    L.ismcts_launch_count.restype = u64
    L.ismcts_launch_count.argtypes = [vp]
    L.ismcts_version.restype = C.c_char_p
    _lib = L
    return L


class EngineError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"ismcts error {code}: {text}")
        self.code = code


def pack_states(states) -> np.ndarray:
    """Packs a sequence of KwState / MnkState into one contiguous uint8 array."""
    size = C.sizeof(type(states[0]))
    buf = np.empty(len(states) * size, dtype=np.uint8)
    for i, s in enumerate(states):
        buf[i * size:(i + 1) * size] = np.frombuffer(bytes(s), dtype=np.uint8)
    return buf


class Engine:
    """One search context on one CUDA device (ismcts_create / ismcts_destroy)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.ismcts_create(device, C.byref(self.ctx))
        if rc != OK:
            raise EngineError(rc, self.lib.ismcts_last_error(None).decode())
        self.device = device

    def close(self):
        if self.ctx:
            self.lib.ismcts_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != OK:
            raise EngineError(rc, self.lib.ismcts_last_error(self.ctx).decode())

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.ismcts_set_stream(self.ctx, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._check(self.lib.ismcts_synchronize(self.ctx))

    @property
    def launch_count(self) -> int:
        return int(self.lib.ismcts_launch_count(self.ctx))

    def make_desc(self, game, n_positions, trees, iterations=0, time_ns=0, seed=1, solver=SOLVER_SO,
                  policy=POLICY_UCB1, exploration=0.7, trees_total=None, tree_base=0, pos_base=0,
                  dump_tree=NO_DUMP, flags=0) -> SearchDesc:
        return SearchDesc(game, solver, policy, n_positions, trees, trees_total or (tree_base + trees), tree_base,
                          pos_base, iterations, time_ns, seed, exploration, dump_tree, flags)

    def search(self, desc: SearchDesc, roots: np.ndarray, dump_capacity: int = 0) -> dict:
        """Host-buffer search (ismcts_search).  `roots` is the packed uint8 array of positions."""
        stride = self.lib.ismcts_move_stride(desc.game)
        n = desc.n_positions
        res = {
            "visits": np.zeros((n, stride), dtype=np.uint64),
            "score2": np.zeros((n, stride), dtype=np.uint64),
            "avail": np.zeros((n, stride), dtype=np.uint64),
            "iterations_done": np.zeros((n, desc.trees), dtype=np.uint64),
            "best_move": np.zeros(n, dtype=np.uint32),
        }
        nodes = np.zeros(max(dump_capacity, 1), dtype=NODE_DTYPE)
        out = SearchOut(res["visits"].ctypes.data, res["score2"].ctypes.data, res["avail"].ctypes.data,
                        res["iterations_done"].ctypes.data, res["best_move"].ctypes.data,
                        nodes.ctypes.data if dump_capacity else None, dump_capacity, 0)
        roots = np.ascontiguousarray(roots)
        self._check(self.lib.ismcts_search(self.ctx, C.byref(desc), roots.ctypes.data, C.byref(out)))
        if dump_capacity:
            res["nodes"] = nodes[:out.dump_count].copy()
        return res

    def search_multi(self, others: list, desc: SearchDesc, roots: np.ndarray) -> dict:
        """ismcts_search_multi over this engine and `others` (engines on other devices)."""
        stride = self.lib.ismcts_move_stride(desc.game)
        n = desc.n_positions
        res = {"visits": np.zeros((n, stride), dtype=np.uint64), "score2": np.zeros((n, stride), dtype=np.uint64),
               "avail": np.zeros((n, stride), dtype=np.uint64), "iterations_done": np.zeros((n, desc.trees), dtype=np.uint64),
               "best_move": np.zeros(n, dtype=np.uint32)}
        out = SearchOut(res["visits"].ctypes.data, res["score2"].ctypes.data, res["avail"].ctypes.data,
                        res["iterations_done"].ctypes.data, res["best_move"].ctypes.data, None, 0, 0)
        ctxs = (C.c_void_p * (1 + len(others)))(self.ctx, *[e.ctx for e in others])
        roots = np.ascontiguousarray(roots)
        self._check(self.lib.ismcts_search_multi(ctxs, 1 + len(others), C.byref(desc), roots.ctypes.data, C.byref(out)))
        return res

    def search_device(self, desc: SearchDesc, d_roots: int, d_visits: int, d_score2: int, d_avail: int,
                      d_iters: int):
        """Device-resident asynchronous search (ismcts_search_device); arguments are device addresses."""
        self._check(self.lib.ismcts_search_device(self.ctx, C.byref(desc), C.c_void_p(d_roots),
                                                  C.c_void_p(d_visits), C.c_void_p(d_score2),
                                                  C.c_void_p(d_avail), C.c_void_p(d_iters)))

    def dump_tree(self, index: int, capacity: int, player: int | None = None) -> np.ndarray:
        """Node records of tree `index` of the last search (MO: the tree of `player`, default the
        root's player to move), in creation order."""
        nodes = np.zeros(capacity, dtype=NODE_DTYPE)
        n = C.c_uint32(0)
        if player is None:
            self._check(self.lib.ismcts_dump_tree(self.ctx, index, nodes.ctypes.data, capacity, C.byref(n)))
        else:
            self._check(self.lib.ismcts_dump_tree_of(self.ctx, index, player, nodes.ctypes.data, capacity, C.byref(n)))
        return nodes[:n.value].copy()

    def playouts(self, game: int, roots: np.ndarray, n_positions: int, per_position: int, seed: int) -> np.ndarray:
        out = np.zeros((n_positions, per_position), dtype=np.uint32)
        roots = np.ascontiguousarray(roots)
        self._check(self.lib.ismcts_playouts(self.ctx, game, roots.ctypes.data, n_positions, per_position, seed,
                                             out.ctypes.data))
        return out


# ---- host game helpers (no device needed) ----

def kw_deal(seed: int, stream: int, players: int = 4) -> KwState:
    s = KwState()
    rc = load_library().ismcts_kw_deal(seed, stream, players, C.byref(s))
    if rc != OK:
        raise EngineError(rc, "ismcts_kw_deal")
    return s


def kw_legal_mask(s: KwState) -> int:
    return int(load_library().ismcts_kw_legal(C.byref(s)))


def kw_apply(s: KwState, move: int, seed: int = 0, stream: int = 0, counter: list | None = None) -> int:
    d = C.c_uint32(counter[0] if counter else 0)
    rc = load_library().ismcts_kw_apply(C.byref(s), move, seed, stream, C.byref(d))
    if counter is not None:
        counter[0] = d.value
    return rc


def mnk_init(m: int, n: int, k: int) -> MnkState:
    s = MnkState()
    rc = load_library().ismcts_mnk_init(m, n, k, C.byref(s))
    if rc != OK:
        raise EngineError(rc, "ismcts_mnk_init")
    return s


def mnk_legal(s: MnkState) -> list[int]:
    legal = (C.c_uint64 * 4)()
    load_library().ismcts_mnk_legal(C.byref(s), legal)
    return [c for c in range(256) if legal[c >> 6] >> (c & 63) & 1]


def mnk_apply(s: MnkState, move: int) -> int:
    return load_library().ismcts_mnk_apply(C.byref(s), move)


def goof_init(seed: int, stream: int = 0) -> GoofState:
    s = GoofState()
    rc = load_library().ismcts_goof_init(seed, stream, C.byref(s))
    if rc != OK:
        raise EngineError(rc, "ismcts_goof_init")
    return s


def goof_legal(s: GoofState) -> list[int]:
    mask = int(load_library().ismcts_goof_legal(C.byref(s)))
    return [c for c in range(64) if mask >> c & 1]


def goof_apply(s: GoofState, move: int) -> int:
    return load_library().ismcts_goof_apply(C.byref(s), move)


def phantom_init(m: int = 3, n: int = 3, k: int = 3) -> PhantomState:
    s = PhantomState()
    rc = load_library().ismcts_phantom_init(m, n, k, C.byref(s))
    if rc != OK:
        raise EngineError(rc, "ismcts_phantom_init")
    return s


def phantom_legal(s: PhantomState) -> list[int]:
    legal = (C.c_uint64 * 4)()
    load_library().ismcts_phantom_legal(C.byref(s), legal)
    return [c for c in range(256) if legal[c >> 6] >> (c & 63) & 1]


def phantom_apply(s: PhantomState, move: int) -> int:
    return load_library().ismcts_phantom_apply(C.byref(s), move)
