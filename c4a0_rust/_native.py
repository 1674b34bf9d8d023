"""Python-facing classes of the reference's native module, backed by libc4a0_engine.so.

Mirrors rust/src/pybridge.rs, rust/src/types.rs and rust/src/lib.rs:
  GameMetadata      types.rs:36-60      GameResult   types.rs:62-100
  Sample            types.rs:102-153    PlayGamesResult  pybridge.rs:55-158
  play_games        pybridge.rs:20-53   run_tui      pybridge.rs:231-251 (not supported here)

Results live as struct-of-arrays (what the GPU produced) and the per-sample Python objects the
reference API exposes are created lazily.  Error behaviour: where the reference panics
(`PanicException`) this raises TypeError / ValueError / RuntimeError instead; `from_cbor`,
`__setstate__` raise ValueError like the reference's `pyify_err`.
"""

from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from c4a0_b200 import _lib as L
from c4a0_b200 import engine as E

from . import _cbor

N_COLS = L.N_COLS
N_ROWS = L.N_ROWS
BUF_N_CHANNELS = L.BUF_N_CHANNELS

_U64 = (1 << 64) - 1


def _u64(name: str, v) -> int:
    if isinstance(v, bool) or not isinstance(v, (int, np.integer)):
        raise TypeError(f"{name} must be an int")
    v = int(v)
    if not 0 <= v <= _U64:
        raise OverflowError(f"{name} does not fit in u64")
    return v


class GameMetadata:
    __slots__ = ("_game_id", "_player0_id", "_player1_id")

    def __init__(self, game_id: int, player0_id: int, player1_id: int) -> None:
        # plain ints in range (what a trainer passes, training.py:181) skip the per-field checks: a generation
        # builds one object per game and 16,384 of them cost 28 ms the slow way
        if (type(game_id) is int and type(player0_id) is int and type(player1_id) is int
                and 0 <= game_id <= _U64 and 0 <= player0_id <= _U64 and 0 <= player1_id <= _U64):
            self._game_id, self._player0_id, self._player1_id = game_id, player0_id, player1_id
            return
        self._game_id = _u64("game_id", game_id)
        self._player0_id = _u64("player0_id", player0_id)
        self._player1_id = _u64("player1_id", player1_id)

    game_id = property(lambda self: self._game_id)
    player0_id = property(lambda self: self._player0_id)
    player1_id = property(lambda self: self._player1_id)

    def __repr__(self) -> str:
        return f"GameMetadata(game_id={self._game_id}, player0_id={self._player0_id}, player1_id={self._player1_id})"

    def __eq__(self, other) -> bool:
        return isinstance(other, GameMetadata) and (
            (self._game_id, self._player0_id, self._player1_id) == (other._game_id, other._player0_id, other._player1_id)
        )

    def __hash__(self):
        return hash((self._game_id, self._player0_id, self._player1_id))


def _planes(mask: int, value: int) -> np.ndarray:
    """c4r.rs:378-392: [2,6,7] f32, channel 0 = side to move, channel 1 = opponent."""
    bits = np.arange(42, dtype=np.uint64)
    mine = (np.uint64(mask & value) >> bits) & np.uint64(1)
    theirs = (np.uint64(mask & ~value & ((1 << 42) - 1)) >> bits) & np.uint64(1)
    return np.concatenate([mine, theirs]).astype(np.float32).reshape(2, 6, 7)


class Sample:
    """A training sample; like the reference it exposes no field getters (types.rs:112-153)."""

    __slots__ = ("_mask", "_value", "_policy", "_q_penalty", "_q_no_penalty")

    def __init__(self, mask: int, value: int, policy, q_penalty, q_no_penalty):
        self._mask = int(mask)
        self._value = int(value)
        self._policy = np.array(policy, dtype=np.float32).reshape(7)
        self._q_penalty = np.float32(q_penalty)
        self._q_no_penalty = np.float32(q_no_penalty)

    def flip_h(self) -> "Sample":
        m, v = E.host_flip_h(self._mask, self._value)
        return Sample(m, v, self._policy[::-1].copy(), self._q_penalty, self._q_no_penalty)

    def to_numpy(self):
        return (
            _planes(self._mask, self._value),
            self._policy.copy(),
            np.array(self._q_penalty, dtype=np.float32),
            np.array(self._q_no_penalty, dtype=np.float32),
        )

    def pos_str(self) -> str:
        """c4r.rs:395-413: rows top-down, red = side to move, blue = opponent."""
        rows = []
        for r in range(N_ROWS - 1, -1, -1):
            line = ""
            for c in range(N_COLS):
                bit = 1 << (r * N_COLS + c)
                if not self._mask & bit:
                    line += "⚫"
                elif self._value & bit:
                    line += "\U0001f534"
                else:
                    line += "\U0001f535"
            rows.append(line)
        return "\n".join(rows)

    def _key(self):
        return (self._mask, self._value, self._policy.tobytes(), self._q_penalty.tobytes(), self._q_no_penalty.tobytes())

    def __eq__(self, other) -> bool:
        return isinstance(other, Sample) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self) -> str:
        return f"Sample(mask={self._mask:#x}, value={self._value:#x}, q_penalty={float(self._q_penalty):.4f})"


class GameResult:
    __slots__ = ("_metadata", "_samples")

    def __init__(self, metadata: GameMetadata, samples: List[Sample]):
        self._metadata = metadata
        self._samples = samples

    @property
    def metadata(self) -> GameMetadata:
        return self._metadata

    @property
    def samples(self) -> List[Sample]:
        return list(self._samples)

    def player0_score(self) -> float:
        """types.rs:73-100: 1 / 0 / 0.5 from the first terminal sample, flipped at odd ply."""
        for s in self._samples:
            t = E.host_terminal_state(s._mask, s._value)
            if t:
                score = {1: 1.0, 2: 0.0, 3: 0.5}[t]
                return 1.0 - score if bin(s._mask).count("1") % 2 == 1 else score
        raise RuntimeError("player0_score called on an unfinished game")


class PlayGamesResult:
    """pybridge.rs:55-158.  Pickles as CBOR bytes, like the reference."""

    def __init__(self) -> None:
        self._meta = np.zeros((0, 3), np.uint64)
        self._soa: Optional[E.GameSamples] = E.GameSamples(
            np.zeros(0, np.uint32), np.zeros((0, 43), np.uint64), np.zeros((0, 43), np.uint64),
            np.zeros((0, 43, 7), np.float32), np.zeros((0, 43), np.float32), np.zeros((0, 43), np.float32),
        )
        self._results: Optional[List[GameResult]] = None
        self._run_info = None

    @classmethod
    def _from_soa(cls, meta: np.ndarray, soa: E.GameSamples) -> "PlayGamesResult":
        r = cls()
        r._meta = np.ascontiguousarray(meta, dtype=np.uint64).reshape(-1, 3)
        r._soa = soa
        return r

    @classmethod
    def _from_results(cls, results: List[GameResult]) -> "PlayGamesResult":
        n = len(results)
        soa = E.GameSamples(
            np.zeros(n, np.uint32), np.zeros((n, 43), np.uint64), np.zeros((n, 43), np.uint64),
            np.zeros((n, 43, 7), np.float32), np.zeros((n, 43), np.float32), np.zeros((n, 43), np.float32),
        )
        meta = np.zeros((n, 3), np.uint64)
        for i, g in enumerate(results):
            md = g._metadata
            meta[i] = (md.game_id, md.player0_id, md.player1_id)
            if len(g._samples) > 43:
                raise ValueError("a game cannot have more than 43 samples")
            soa.n_samples[i] = len(g._samples)
            for k, s in enumerate(g._samples):
                soa.mask[i, k], soa.value[i, k] = s._mask, s._value
                soa.policy[i, k] = s._policy
                soa.q_penalty[i, k], soa.q_no_penalty[i, k] = s._q_penalty, s._q_no_penalty
        r = cls._from_soa(meta, soa)
        r._results = list(results)
        return r

    # ---- reference API -----------------------------------------------------------------------
    @property
    def results(self) -> List[GameResult]:
        if self._results is None:
            # bulk conversions first (one C loop per array), then plain slot assignments per sample: the per-
            # sample cost is object creation only (~1 us), not numpy scalar traffic
            soa, out = self._soa, []
            counts = soa.n_samples.tolist()
            meta = self._meta.tolist()
            valid = np.arange(43)[None, :] < soa.n_samples[:, None]
            masks, values = soa.mask[valid].tolist(), soa.value[valid].tolist()
            pol = np.ascontiguousarray(soa.policy[valid])
            qps, qns = soa.q_penalty[valid], soa.q_no_penalty[valid]
            new, j = Sample.__new__, 0
            for i, n in enumerate(counts):
                samples = []
                for _ in range(n):
                    s = new(Sample)
                    s._mask, s._value, s._policy, s._q_penalty, s._q_no_penalty = masks[j], values[j], pol[j], qps[j], qns[j]
                    samples.append(s)
                    j += 1
                md = new(GameMetadata)
                md._game_id, md._player0_id, md._player1_id = meta[i]
                out.append(GameResult(md, samples))
            self._results = out
        return list(self._results)

    def to_cbor(self) -> bytes:
        """serde_cbor's encoding of the result (pybridge.rs:73-81), written by the library's bulk encoder
        (csrc/cbor.cu); c4a0_rust/_cbor.py is the same codec in Python."""
        import ctypes as C

        soa = self._soa
        arrs = [np.ascontiguousarray(self._meta, dtype=np.uint64), np.ascontiguousarray(soa.n_samples, dtype=np.uint32),
                np.ascontiguousarray(soa.mask, dtype=np.uint64), np.ascontiguousarray(soa.value, dtype=np.uint64),
                np.ascontiguousarray(soa.policy, dtype=np.float32), np.ascontiguousarray(soa.q_penalty, dtype=np.float32),
                np.ascontiguousarray(soa.q_no_penalty, dtype=np.float32)]
        n, need = len(arrs[1]), C.c_size_t(0)
        lib = L.lib()
        L.check(lib.c4a0_results_to_cbor(*[L.ptr(a) for a in arrs], n, None, 0, C.byref(need)))
        out = np.empty(need.value, np.uint8)
        L.check(lib.c4a0_results_to_cbor(*[L.ptr(a) for a in arrs], n, L.ptr(out), out.size, C.byref(need)))
        return out.tobytes()

    @staticmethod
    def from_cbor(cbor: bytes) -> "PlayGamesResult":
        import ctypes as C

        buf = np.frombuffer(bytes(cbor), dtype=np.uint8)
        lib, n = L.lib(), C.c_uint32(0)
        if lib.c4a0_results_from_cbor(L.ptr(buf), buf.size, C.byref(n), None, None, None, None, None, None, None) != 0:
            raise ValueError((lib.c4a0_last_error() or b"invalid PlayGamesResult CBOR").decode("utf-8", "replace"))
        g = n.value
        meta = np.zeros((g, 3), np.uint64)
        soa = E.GameSamples(np.zeros(g, np.uint32), np.zeros((g, 43), np.uint64), np.zeros((g, 43), np.uint64),
                            np.zeros((g, 43, 7), np.float32), np.zeros((g, 43), np.float32), np.zeros((g, 43), np.float32))
        if lib.c4a0_results_from_cbor(L.ptr(buf), buf.size, C.byref(n), L.ptr(meta), L.ptr(soa.n_samples), L.ptr(soa.mask),
                                      L.ptr(soa.value), L.ptr(soa.policy), L.ptr(soa.q_penalty), L.ptr(soa.q_no_penalty)) != 0:
            raise ValueError((lib.c4a0_last_error() or b"invalid PlayGamesResult CBOR").decode("utf-8", "replace"))
        return PlayGamesResult._from_soa(meta, soa)

    def __getstate__(self) -> bytes:
        return self.to_cbor()

    def __setstate__(self, state: bytes) -> None:
        other = PlayGamesResult.from_cbor(state)
        self._meta, self._soa, self._results = other._meta, other._soa, other._results
        self._run_info = None  # pickle bypasses __init__

    def __add__(self, other: "PlayGamesResult") -> "PlayGamesResult":
        if not isinstance(other, PlayGamesResult):
            raise TypeError("can only add PlayGamesResult to PlayGamesResult")
        a, b = self._soa, other._soa
        soa = E.GameSamples(*[np.concatenate([getattr(a, f), getattr(b, f)]) for f in
                              ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty")])
        return PlayGamesResult._from_soa(np.concatenate([self._meta, other._meta]), soa)

    def split_train_test(self, train_frac: float, seed: int) -> Tuple[List[Sample], List[Sample]]:
        """pybridge.rs:107-120: shuffle whole games with StdRng(seed); the first
        round(len * train_frac) games are the training set."""
        games = self.results
        order = E.host_shuffle(_u64("seed", seed), len(games))
        x = np.float32(len(games)) * np.float32(train_frac)
        n_train = int(np.floor(np.abs(x) + np.float32(0.5)))  # f32::round: half away from zero
        n_train = max(0, min(len(games), n_train))
        shuffled = [games[i] for i in order]
        train = [s for g in shuffled[:n_train] for s in g._samples]
        test = [s for g in shuffled[n_train:] for s in g._samples]
        return train, test

    def score_policies(self, solver_path: str, solver_book_path: str, solution_cache_path: str) -> float:
        raise NotImplementedError(
            "score_policies drives an external solver binary (rust/src/solver.rs); it is outside the self-play hot path"
        )

    def unique_positions(self) -> int:
        soa = self._soa
        valid = np.arange(43)[None, :] < soa.n_samples[:, None]
        keys = np.stack([soa.mask[valid], soa.value[valid]], axis=1)
        return int(np.unique(keys, axis=0).shape[0]) if keys.size else 0

    # ---- additive, not in the reference: bulk export for the trainer --------
    def to_arrays(self):
        """(pos f32[S,2,6,7], policy f32[S,7], q_penalty f32[S], q_no_penalty f32[S]) of all samples."""
        soa = self._soa
        valid = np.arange(43)[None, :] < soa.n_samples[:, None]
        mask, value = soa.mask[valid], soa.value[valid]
        bits = np.arange(42, dtype=np.uint64)[None, :]
        mine = ((mask & value)[:, None] >> bits) & np.uint64(1)
        theirs = ((mask & ~value)[:, None] >> bits) & np.uint64(1)
        pos = np.concatenate([mine, theirs], axis=1).astype(np.float32).reshape(-1, 2, 6, 7)
        return pos, soa.policy[valid], soa.q_penalty[valid], soa.q_no_penalty[valid]


for _cls in (GameMetadata, Sample, GameResult, PlayGamesResult):
    _cls.__module__ = "c4a0_rust"  # pybridge.rs:57-59: pickles name the class as c4a0_rust.<Class>


_SESSION = {"key": None, "sess": None}


def _session(n_slots, n_req, n_iter, c_expl, c_pen, dtype, device, stride, n_lanes, offset=0, eval_cache=False,
             spec_rows=0):
    """One engine (tree arenas, NN I/O tensors, captured graphs) is kept between calls with the same
    configuration — a training loop calls play_games once per generation with identical settings."""
    from c4a0_b200 import selfplay
    from c4a0_b200.selfplay import SelfPlaySession

    knobs = tuple(sorted((k, v) for k, v in selfplay.DEFAULTS.items() if k in ("n_lanes", "dedup", "max_inline_sims", "arena_blocks", "eval_cache_entries", "speculate", "spec_rows", "dirichlet")))
    key = (n_slots, n_iter, c_expl, c_pen, dtype, device, stride, offset, n_lanes, knobs, eval_cache, spec_rows)
    if _SESSION["key"] == key and _SESSION["sess"] is not None and _SESSION["cap"] >= n_req:
        return _SESSION["sess"]
    close_cached_session()
    sess = SelfPlaySession(n_slots, n_req, n_iter, c_expl, c_pen, plane_dtype=dtype, device=device,
                           plane_stride=stride, plane_offset=offset, n_lanes=n_lanes, eval_cache=eval_cache,
                           eval_cache_entries=selfplay.DEFAULTS["eval_cache_entries"],
                           speculate=bool(eval_cache and spec_rows > 0), spec_rows=spec_rows)
    _SESSION.update(key=key, sess=sess, cap=n_req)
    return sess


_MODULE_EVALUATORS = {}  # id(module) -> (weakref to the module, its evaluator)


def _module_evaluator(module):
    """An nn.Module handed to play_games: evaluate it in its own precision (bf16 parameters -> the library's
    tcgen05 kernel, float32 -> PyTorch in float32, the reference's precision).  The evaluator is kept per
    module, so a loop that calls play_games with the same module re-loads its weights in place instead of
    re-folding into new buffers and re-capturing CUDA graphs; the module's training flag is left as found."""
    import weakref

    from c4a0_b200.selfplay import DeviceEvaluator

    key = id(module)
    hit = _MODULE_EVALUATORS.get(key)
    prev = hit[1] if hit is not None and hit[0]() is module else None
    was_training = module.training
    p = next(module.parameters(), None)
    try:
        ev = DeviceEvaluator.from_model(module, p.dtype if p is not None else __import__("torch").float32, reuse=prev)
    finally:
        module.train(was_training)
    for k in [k for k, (ref, _) in _MODULE_EVALUATORS.items() if ref() is None]:
        del _MODULE_EVALUATORS[k]
    _MODULE_EVALUATORS[key] = (weakref.ref(module), ev)
    return ev


def close_cached_session() -> None:
    """Free the engine kept by the last play_games call (device memory is released)."""
    if _SESSION["sess"] is not None:
        _SESSION["sess"].close()
    _SESSION.update(key=None, sess=None, cap=0)


def play_games(
    reqs: Sequence[GameMetadata],
    max_nn_batch_size: int,
    n_mcts_iterations: int,
    c_exploration: float,
    c_ply_penalty: float,
    py_eval_pos_cb: Callable,
) -> PlayGamesResult:
    """Play the requested games with MCTS self-play on the GPU (pybridge.rs:20-53).

    `py_eval_pos_cb` is either the reference's numpy callback `cb(model_id, ndarray[B,2,6,7]) ->
    (policy[B,7], q_penalty[B], q_no_penalty[B])`, or — the fast path — a
    `c4a0_b200.selfplay.DeviceEvaluator` / `torch.nn.Module` that takes the planes as a CUDA tensor
    and returns CUDA tensors, in which case nothing crosses PCIe during the search.  Results are
    returned in request order (the reference returns completion order, which is nondeterministic).
    """
    import torch

    from c4a0_b200 import selfplay
    from c4a0_b200.native_net import NativeEvaluator
    from c4a0_b200.selfplay import BuiltinEvaluator, DeviceEvaluator, MultiModelEvaluator

    reqs = list(reqs)
    if not all(type(r) is GameMetadata or isinstance(r, GameMetadata) for r in reqs):
        raise TypeError("reqs must be a list of GameMetadata")
    if int(max_nn_batch_size) < 1 or int(n_mcts_iterations) < 1:
        raise ValueError("max_nn_batch_size and n_mcts_iterations must be >= 1")
    if not callable(py_eval_pos_cb):
        raise TypeError("py_eval_pos_cb must be callable")
    if not reqs:
        return PlayGamesResult()
    meta = np.empty((len(reqs), 3), dtype=np.uint64)
    meta[:, 0] = [r._game_id for r in reqs]
    meta[:, 1] = [r._player0_id for r in reqs]
    meta[:, 2] = [r._player1_id for r in reqs]
    n_slots = min(len(reqs), int(max_nn_batch_size))
    fast = isinstance(py_eval_pos_cb, (DeviceEvaluator, MultiModelEvaluator, BuiltinEvaluator, NativeEvaluator, torch.nn.Module))
    if fast:
        if isinstance(py_eval_pos_cb, torch.nn.Module):
            py_eval_pos_cb = _module_evaluator(py_eval_pos_cb)
        ids = set(np.unique(meta[:, 1:]).tolist())
        if isinstance(py_eval_pos_cb, MultiModelEvaluator):
            missing = ids - set(py_eval_pos_cb.evaluators)
            if missing:
                raise ValueError(f"no evaluator for model ids {sorted(missing)}")
        elif len(ids) != 1 and not isinstance(py_eval_pos_cb, BuiltinEvaluator):
            raise ValueError("one device evaluator plays one model against itself; pass a MultiModelEvaluator "
                             "{model_id: evaluator} (or the numpy callback) for tournaments")
        # max_nn_batch_size bounds every network batch (pybridge.rs:20-53, self_play.rs:216-220).  The
        # speculative rows (children of expanded leaves evaluated in the spare rows of a small batch) only
        # use what the caller's bound leaves above the resident games: n_slots + spec_rows <= max_nn_batch_size.
        use_cache = bool(selfplay.DEFAULTS["eval_cache"])
        spec_rows = 0
        if use_cache and selfplay.DEFAULTS["speculate"]:
            spec_rows = min(selfplay.DEFAULTS["spec_rows"] or 8192, n_slots, int(max_nn_batch_size) - n_slots)
            if spec_rows < 32:
                spec_rows = 0
        sess = _session(
            n_slots, len(reqs), int(n_mcts_iterations), float(c_exploration), float(c_ply_penalty),
            py_eval_pos_cb.dtype, torch.cuda.current_device(), py_eval_pos_cb.plane_stride, None,
            getattr(py_eval_pos_cb, "plane_offset", 0), use_cache, spec_rows,
        )
        try:
            soa, info = sess.play(meta[:, 0], meta[:, 1], meta[:, 2], py_eval_pos_cb, fetch=bool(selfplay.DEFAULTS.get("fetch", True)))
        except Exception:
            close_cached_session()  # never keep an engine in an unknown state
            raise
        if soa is None:
            # selfplay.DEFAULTS["fetch"] = False (additive): the samples stay in the engine's device store for
            # dist.gather_session_samples() / SelfPlaySession.export_tensors(); the result object is empty
            out = PlayGamesResult()
            out._run_info = info
            return out
    else:
        sess = _session(
            n_slots, len(reqs), int(n_mcts_iterations), float(c_exploration), float(c_ply_penalty),
            torch.float32, torch.cuda.current_device(), 84, 1, 0,
        )
        try:
            soa, info = sess.play_callback(meta[:, 0], meta[:, 1], meta[:, 2], py_eval_pos_cb, int(max_nn_batch_size))
        except Exception:
            close_cached_session()
            raise
    out = PlayGamesResult._from_soa(meta, soa)
    out._run_info = info  # additive: counters and timings of this call (c4a0_b200.selfplay.RunInfo)
    return out


def run_tui(py_eval_pos_cb, max_mcts_iters: int, c_exploration: float, c_ply_penalty: float) -> None:
    raise NotImplementedError("the terminal UI (rust/src/tui.rs) is outside the self-play hot path")
