"""ctypes binding of the CPU ORACLE (oracle/c4a0_oracle.c, oracle/selfplay_threads.cpp).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from c4a0_b200/ or c4a0_rust/.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libc4a0_oracle.so")

N_ROWS, N_COLS, BUF_LEN, MAX_SAMPLES = 6, 7, 84, 43
NONE, PLAYER_WIN, OPPONENT_WIN, DRAW = 0, 1, 2, 3


def build(force: bool = False) -> str:
    """Compile the oracle with oracle/Makefile (gcc/g++ only)."""
    srcs = [os.path.join(_HERE, f) for f in ("c4a0_oracle.c", "c4a0_oracle.h", "selfplay_threads.cpp", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if stale:
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _LIB_PATH


class Pos(C.Structure):
    _fields_ = [("mask", C.c_uint64), ("value", C.c_uint64)]

    def key(self) -> Tuple[int, int]:
        return (int(self.mask), int(self.value))

    def __repr__(self) -> str:
        return f"Pos(mask={self.mask:#x}, value={self.value:#x})"


class Sample(C.Structure):
    _fields_ = [("pos", Pos), ("policy", C.c_float * 7), ("q_penalty", C.c_float), ("q_no_penalty", C.c_float)]


class Metadata(C.Structure):
    _fields_ = [("game_id", C.c_uint64), ("player0_id", C.c_uint64), ("player1_id", C.c_uint64)]


class Stats(C.Structure):
    _fields_ = [
        ("sims", C.c_uint64),
        ("nn_evals", C.c_uint64),
        ("terminal_leaf_sims", C.c_uint64),
        ("terminal_root_sims", C.c_uint64),
        ("moves", C.c_uint64),
        ("samples", C.c_uint64),
        ("select_depth_sum", C.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


EVAL_FN = C.CFUNCTYPE(
    None, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Pos), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)
)

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    F7 = C.c_float * 7
    L.c4o_make_move.argtypes = [Pos, C.c_int, C.POINTER(Pos)]
    L.c4o_make_move.restype = C.c_int
    L.c4o_get.argtypes = [Pos, C.c_int, C.c_int]
    L.c4o_get.restype = C.c_int
    L.c4o_ply.argtypes = [Pos]
    L.c4o_ply.restype = C.c_int
    L.c4o_invert.argtypes = [Pos]
    L.c4o_invert.restype = Pos
    L.c4o_terminal_state.argtypes = [Pos]
    L.c4o_terminal_state.restype = C.c_int
    L.c4o_terminal_value.argtypes = [Pos, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.c4o_terminal_value.restype = C.c_int
    L.c4o_legal_moves.argtypes = [Pos]
    L.c4o_legal_moves.restype = C.c_uint
    L.c4o_flip_h.argtypes = [Pos]
    L.c4o_flip_h.restype = Pos
    L.c4o_write_planes.argtypes = [Pos, C.POINTER(C.c_float)]
    L.c4o_win_masks.restype = C.POINTER(C.c_uint64)
    L.c4o_from_moves.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(Pos)]
    L.c4o_from_moves.restype = C.c_int
    L.c4o_from_str.argtypes = [C.c_char_p, C.POINTER(Pos)]
    L.c4o_from_str.restype = C.c_int
    L.c4o_to_str.argtypes = [Pos, C.c_char_p, C.c_size_t]
    L.c4o_to_str.restype = C.c_int
    L.c4o_random_pos.argtypes = [C.POINTER(C.c_uint8), C.c_int]
    L.c4o_random_pos.restype = Pos
    L.c4o_softmax.argtypes = [F7, F7]
    L.c4o_softmax.restype = C.c_int
    L.c4o_apply_temperature.argtypes = [F7, C.c_float, F7]
    L.c4o_seed_from_u64.argtypes = [C.c_uint64, C.c_uint32 * 8]
    L.c4o_chacha_block.argtypes = [C.c_uint32 * 8, C.c_uint64, C.c_int, C.c_uint32 * 16]
    L.c4o_chacha_block_nonce.argtypes = [C.c_uint32 * 8, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32 * 16]
    L.c4o_weighted_index_sample.argtypes = [F7, C.c_uint64]
    L.c4o_weighted_index_sample.restype = C.c_int
    L.c4o_shuffle_indices.argtypes = [C.c_uint64, C.POINTER(C.c_uint32), C.c_size_t]
    L.c4o_game_new.argtypes = [Pos, Metadata]
    L.c4o_game_new.restype = C.c_void_p
    L.c4o_game_free.argtypes = [C.c_void_p]
    L.c4o_game_root_pos.argtypes = [C.c_void_p]
    L.c4o_game_root_pos.restype = Pos
    L.c4o_game_leaf_pos.argtypes = [C.c_void_p]
    L.c4o_game_leaf_pos.restype = Pos
    L.c4o_game_leaf_model_id.argtypes = [C.c_void_p]
    L.c4o_game_leaf_model_id.restype = C.c_uint64
    L.c4o_game_on_received_policy.argtypes = [C.c_void_p, F7, C.c_float, C.c_float, C.c_float, C.c_float]
    L.c4o_game_make_move.argtypes = [C.c_void_p, C.c_int, C.c_float]
    L.c4o_game_make_move.restype = C.c_int
    L.c4o_game_make_random_move.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.c4o_game_make_random_move.restype = C.c_int
    L.c4o_game_root_visit_count.argtypes = [C.c_void_p]
    L.c4o_game_root_visit_count.restype = C.c_uint64
    L.c4o_game_root_policy.argtypes = [C.c_void_p, F7]
    L.c4o_game_root_q_penalty.argtypes = [C.c_void_p]
    L.c4o_game_root_q_penalty.restype = C.c_float
    L.c4o_game_root_q_no_penalty.argtypes = [C.c_void_p]
    L.c4o_game_root_q_no_penalty.restype = C.c_float
    L.c4o_game_n_moves.argtypes = [C.c_void_p]
    L.c4o_game_n_moves.restype = C.c_int
    L.c4o_game_to_result.argtypes = [C.c_void_p, C.c_float, C.POINTER(Sample)]
    L.c4o_game_to_result.restype = C.c_int
    L.c4o_game_dump_tree.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_size_t]
    L.c4o_game_dump_tree.restype = C.c_size_t
    sp_args = [
        C.POINTER(Metadata), C.c_size_t, C.c_int, C.c_uint64, C.c_float, C.c_float, C.c_void_p, C.c_void_p,
    ]
    L.c4o_self_play.argtypes = sp_args + [C.POINTER(Sample), C.POINTER(C.c_int), C.POINTER(Stats)]
    L.c4o_self_play.restype = C.c_int
    L.c4o_self_play_threaded.argtypes = sp_args + [
        C.c_int, C.POINTER(Sample), C.POINTER(C.c_int), C.POINTER(Stats), C.POINTER(C.c_uint64),
    ]
    L.c4o_self_play_threaded.restype = C.c_int
    L.c4o_self_play_threaded_budget.argtypes = sp_args + [
        C.c_int, C.c_uint64, C.POINTER(Sample), C.POINTER(C.c_int), C.POINTER(Stats), C.POINTER(C.c_uint64),
    ]
    L.c4o_self_play_threaded_budget.restype = C.c_int
    L.c4o_player0_score.argtypes = [C.POINTER(Sample), C.c_int]
    L.c4o_player0_score.restype = C.c_float
    for ev in (L.c4o_eval_uniform, L.c4o_eval_hash, L.c4o_eval_hash_flat):
        ev.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(Pos), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        ev.restype = None
    L.c4o_logf_array.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.c4o_expf_array.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    _lib = L
    return L


# ----------------------------------------------------------------------------- rules helpers
def make_move(p: Pos, col: int) -> Optional[Pos]:
    out = Pos()
    return out if lib().c4o_make_move(p, col, C.byref(out)) else None


def from_moves(moves: Sequence[int]) -> Pos:
    arr = (C.c_int * len(moves))(*moves)
    out = Pos()
    if not lib().c4o_from_moves(arr, len(moves), C.byref(out)):
        raise ValueError("illegal move sequence")
    return out


def from_str(s: str) -> Pos:
    out = Pos()
    lib().c4o_from_str(s.encode("utf-8"), C.byref(out))
    return out


def to_str(p: Pos) -> str:
    buf = C.create_string_buffer(512)
    n = lib().c4o_to_str(p, buf, 512)
    return buf.raw[:n].decode("utf-8")


def terminal_state(p: Pos) -> int:
    return lib().c4o_terminal_state(p)


def terminal_value(p: Pos, c_ply_penalty: float) -> Optional[Tuple[float, float]]:
    qp, qn = C.c_float(), C.c_float()
    t = lib().c4o_terminal_value(p, c_ply_penalty, C.byref(qp), C.byref(qn))
    return (qp.value, qn.value) if t else None


def legal_moves(p: Pos) -> List[bool]:
    m = lib().c4o_legal_moves(p)
    return [bool(m >> c & 1) for c in range(N_COLS)]


def planes(p: Pos) -> np.ndarray:
    out = np.zeros(BUF_LEN, dtype=np.float32)
    lib().c4o_write_planes(p, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out.reshape(2, N_ROWS, N_COLS)


def random_pos(cols: Sequence[int]) -> Pos:
    arr = (C.c_uint8 * len(cols))(*[int(c) for c in cols])
    return lib().c4o_random_pos(arr, len(cols))


def win_masks() -> List[int]:
    p = lib().c4o_win_masks()
    return [int(p[i]) for i in range(69)]


def softmax(x: Sequence[float]) -> Optional[np.ndarray]:
    F7 = C.c_float * 7
    i, o = F7(*x), F7()
    if not lib().c4o_softmax(i, o):
        return None
    return np.array(list(o), dtype=np.float32)


def apply_temperature(p: Sequence[float], t: float) -> np.ndarray:
    F7 = C.c_float * 7
    i, o = F7(*p), F7()
    lib().c4o_apply_temperature(i, t, o)
    return np.array(list(o), dtype=np.float32)


def weighted_index_sample(w: Sequence[float], seed: int) -> int:
    return lib().c4o_weighted_index_sample((C.c_float * 7)(*w), seed)


def shuffle_indices(seed: int, n: int) -> np.ndarray:
    idx = np.arange(n, dtype=np.uint32)
    lib().c4o_shuffle_indices(seed, idx.ctypes.data_as(C.POINTER(C.c_uint32)), n)
    return idx


def logf(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().c4o_logf_array(x.ctypes.data, out.ctypes.data, x.size)
    return out


def expf(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().c4o_expf_array(x.ctypes.data, out.ctypes.data, x.size)
    return out


# ----------------------------------------------------------------------------- MCTS game
class Game:
    """mcts.rs `MctsGame`."""

    def __init__(self, pos: Optional[Pos] = None, game_id: int = 0, player0_id: int = 0, player1_id: int = 0):
        self._h = lib().c4o_game_new(pos or Pos(0, 0), Metadata(game_id, player0_id, player1_id))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().c4o_game_free(self._h)
            self._h = None

    def on_received_policy(self, policy, q_penalty, q_no_penalty, c_exploration, c_ply_penalty):
        lib().c4o_game_on_received_policy(self._h, (C.c_float * 7)(*policy), q_penalty, q_no_penalty, c_exploration, c_ply_penalty)

    def root_pos(self) -> Pos:
        return lib().c4o_game_root_pos(self._h)

    def leaf_pos(self) -> Pos:
        return lib().c4o_game_leaf_pos(self._h)

    def root_visit_count(self) -> int:
        return lib().c4o_game_root_visit_count(self._h)

    def root_policy(self) -> np.ndarray:
        o = (C.c_float * 7)()
        lib().c4o_game_root_policy(self._h, o)
        return np.array(list(o), dtype=np.float32)

    def root_q_penalty(self) -> float:
        return lib().c4o_game_root_q_penalty(self._h)

    def root_q_no_penalty(self) -> float:
        return lib().c4o_game_root_q_no_penalty(self._h)

    def make_move(self, col: int, c_exploration: float) -> bool:
        return bool(lib().c4o_game_make_move(self._h, col, c_exploration))

    def make_random_move(self, c_exploration: float, temperature: float) -> bool:
        return bool(lib().c4o_game_make_random_move(self._h, c_exploration, temperature))

    def n_moves(self) -> int:
        return lib().c4o_game_n_moves(self._h)

    def to_result(self, c_ply_penalty: float) -> List[Sample]:
        buf = (Sample * MAX_SAMPLES)()
        n = lib().c4o_game_to_result(self._h, c_ply_penalty, buf)
        if n < 0:
            raise ValueError("non-terminal game")
        return [buf[i] for i in range(n)]

    def dump_tree(self) -> np.ndarray:
        need = lib().c4o_game_dump_tree(self._h, None, 0)
        out = np.zeros(need, dtype=np.uint32)
        lib().c4o_game_dump_tree(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), need)
        return out


def builtin_eval(name: str, pos: Pos, model_id: int = 0):
    """One position through the 'uniform' (E0) or 'hash' (E1) synthetic evaluator."""
    fn = {"uniform": lib().c4o_eval_uniform, "hash": lib().c4o_eval_hash, "hash_flat": lib().c4o_eval_hash_flat}[name]
    pol = (C.c_float * 7)()
    qp, qn = C.c_float(), C.c_float()
    fn(None, model_id, 1, C.byref(pos), pol, C.byref(qp), C.byref(qn))
    return list(pol), qp.value, qn.value


def run_mcts(pos: Pos, n_iterations: int, c_exploration: float = 4.0, c_ply_penalty: float = 0.01):
    """mcts.rs:469-485 test helper: constant evaluator (uniform policy 1/7 as 'logits', q = 0)."""
    g = Game(pos)
    pol = [1.0 / 7.0] * 7
    h = g._h
    L = lib()
    arr = (C.c_float * 7)(*pol)
    for _ in range(n_iterations):
        L.c4o_game_on_received_policy(h, arr, 0.0, 0.0, c_exploration, c_ply_penalty)
    return g.root_policy(), g.root_q_penalty(), g.root_q_no_penalty(), g


# ----------------------------------------------------------------------------- self-play
@dataclass
class SelfPlayOutput:
    samples: List[List[Sample]]  # per request index
    stats: dict
    nn_batches: int = 0

    def records(self):
        """Per game: list of (mask, value, policy bits tuple, qp bits, qn bits)."""
        out = []
        for g in self.samples:
            out.append(
                [
                    (
                        int(s.pos.mask),
                        int(s.pos.value),
                        tuple(np.array(list(s.policy), dtype=np.float32).view(np.uint32).tolist()),
                        int(np.float32(s.q_penalty).view(np.uint32)),
                        int(np.float32(s.q_no_penalty).view(np.uint32)),
                    )
                    for s in g
                ]
            )
        return out


def _wrap_eval(evaluator) -> Tuple[object, object]:
    """evaluator: 'uniform' | 'hash' | callable(model_id, list[(mask,value)]) -> (policy[n,7], qp[n], qn[n])."""
    L = lib()
    if evaluator == "uniform":
        return C.cast(L.c4o_eval_uniform, C.c_void_p), None
    if evaluator == "hash":
        return C.cast(L.c4o_eval_hash, C.c_void_p), None
    if evaluator == "hash_flat":
        return C.cast(L.c4o_eval_hash_flat, C.c_void_p), None

    def cb(_user, model_id, n, pos, policy, qp, qn):
        keys = [(int(pos[i].mask), int(pos[i].value)) for i in range(n)]
        pol, a, b = evaluator(int(model_id), keys)
        pol = np.ascontiguousarray(pol, dtype=np.float32).reshape(n, 7)
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(n)
        b = np.ascontiguousarray(b, dtype=np.float32).reshape(n)
        C.memmove(policy, pol.ctypes.data, 4 * 7 * n)
        C.memmove(qp, a.ctypes.data, 4 * n)
        C.memmove(qn, b.ctypes.data, 4 * n)

    fn = EVAL_FN(cb)
    return C.cast(fn, C.c_void_p), fn


def self_play(
    reqs: Sequence[Tuple[int, int, int]],
    max_nn_batch_size: int,
    n_mcts_iterations: int,
    c_exploration: float,
    c_ply_penalty: float,
    evaluator="uniform",
    threaded: bool = False,
    n_threads: int = 0,
) -> SelfPlayOutput:
    L = lib()
    n = len(reqs)
    md = (Metadata * max(n, 1))(*[Metadata(*r) for r in reqs])
    out = (Sample * (max(n, 1) * MAX_SAMPLES))()
    out_n = (C.c_int * max(n, 1))()
    st = Stats()
    fn, keep = _wrap_eval(evaluator)
    nb = C.c_uint64(0)
    if threaded:
        rc = L.c4o_self_play_threaded(
            md, n, max_nn_batch_size, n_mcts_iterations, c_exploration, c_ply_penalty, fn, None, n_threads, out, out_n,
            C.byref(st), C.byref(nb),
        )
    else:
        rc = L.c4o_self_play(md, n, max_nn_batch_size, n_mcts_iterations, c_exploration, c_ply_penalty, fn, None, out, out_n, C.byref(st))
    del keep
    if rc != 0:
        raise RuntimeError(f"oracle self-play failed rc={rc}")
    samples = [[out[i * MAX_SAMPLES + k] for k in range(out_n[i])] for i in range(n)]
    return SelfPlayOutput(samples=samples, stats=st.as_dict(), nn_batches=int(nb.value))


def self_play_parallel(reqs, n_mcts_iterations, c_exploration, c_ply_penalty, evaluator="hash", chunk=16, workers=0):
    """Records (SelfPlayOutput.records() form) of many independent games, played by the serial state machine
    on a pool of host threads (ctypes releases the GIL; the built-in evaluators never call back into
    Python).  Per-game results do not depend on scheduling, so chunks can run anywhere."""
    from concurrent.futures import ThreadPoolExecutor

    assert evaluator in ("uniform", "hash", "hash_flat")
    reqs = list(reqs)
    parts = [reqs[i : i + chunk] for i in range(0, len(reqs), chunk)]
    workers = workers or max(1, (os.cpu_count() or 2))

    def run(part):
        return self_play(part, len(part), n_mcts_iterations, c_exploration, c_ply_penalty, evaluator=evaluator).records()

    with ThreadPoolExecutor(max_workers=workers) as ex:
        out = []
        for r in ex.map(run, parts):
            out.extend(r)
    return out


def player0_score(samples: Sequence[Sample]) -> float:
    arr = (Sample * len(samples))(*samples)
    return float(lib().c4o_player0_score(arr, len(samples)))
