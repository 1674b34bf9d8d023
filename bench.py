#!/usr/bin/env python
"""bench.py — self-play positions/s (and MCTS sims/s) of the B200 engine, next to the reference's
CPU self-play architecture timed on the same box.

A *step* is one whole self-play job through the public API: `c4a0_rust.play_games(reqs, ...)` with
HOST request objects in and HOST samples out (BASELINE.json configs[1]: 16,384 lockstep games x 600
MCTS sims/move, default c4a0 ResNet with random init, per GPU).  From the same K steps:
  value        positions/s over the device-timed search region (CUDA events, inputs resident in HBM)
  e2e.value    positions/s over the wall clock of the API calls (H2D of the requests, weight folding, search,
               D2H of the samples; at N > 1 the NCCL weight broadcast and the packed sample gather to rank 0)
  roofline     the tree kernel `k_step` (apply + select): algorithmic bytes per launch / its average
               device time, sampled with CUDA events on ticks inside the timed region; `traffic` = DRAM bytes of
               that kernel's launches in this job from the round's ncu capture (profiles/rNN_kernel_counters.json)
  nn_roofline  the network kernel `k_net2` against the measured dense bf16 peak, with the tensor-pipe utilisation
               ncu measured for it
  cpu_baseline the oracle's threaded restatement of rust/src/self_play.rs on the host cores
               (rank 0, N=1 only), network evaluated through the numpy callback on cuda:0 as the
               reference does

`--impl reference` times that CPU implementation alone (the Rust crate cannot be built in this
image: no cargo/rustc).
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "self-play positions/sec"
UNIT = "positions/s"
C_EXPLORATION, C_PLY_PENALTY = 6.6, 0.01  # src/c4a0/main.py:42-43


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--games", type=int, default=16384, help="lockstep games per GPU")
    ap.add_argument("--sims", type=int, default=600, help="MCTS iterations per move")
    ap.add_argument("--width", type=int, default=32, help="conv_filter_size of the ResNet")
    ap.add_argument("--nn-dtype", choices=["bf16", "f32"], default="bf16")
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="CPU-baseline budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-nn-device", choices=["auto", "cuda", "cpu"], default="auto",
                    help="where the CPU reference evaluates its network: auto/cuda = through the numpy callback on "
                         "cuda:0, the way the reference is deployed (SURVEY 8d placement i); cpu = host cores only (ii)")
    ap.add_argument("--sample-kernels-every", type=int, default=53)
    ap.add_argument("--lanes", type=int, default=1, help="engines per GPU (2 = tree ticks overlap the other half's network)")
    ap.add_argument("--no-dedup", action="store_true", help="evaluate duplicate leaf positions separately")
    ap.add_argument("--no-eval-cache", action="store_true",
                    help="send every leaf to the network even if this job has evaluated the position before")
    ap.add_argument("--no-speculate", action="store_true", help="do not top small batches up with children of expanded leaves")
    ap.add_argument("--spec-rows", type=int, default=0, help="rows a small batch is topped up to (0 = engine default)")
    ap.add_argument("--eval-cache-entries", type=int, default=0, help="entries of the evaluation cache (0 = engine default)")
    ap.add_argument("--arena-mult", type=float, default=0.0,
                    help="tree blocks per arena half per game as a multiple of the minimum (sims + 2); 0 = session default")
    ap.add_argument("--no-ablation", action="store_true", help="skip the extra step without the evaluation cache")
    ap.add_argument("--max-inline", type=int, default=0, help="terminal-leaf sims per game per tick (0 = engine default)")
    ap.add_argument("--no-fold", action="store_true", help="run the module form of the network instead of the GEMM-folded form")
    ap.add_argument("--host-loop", choices=["native", "python"], default="native",
                    help="python = eager network + per-tick polling (what ncu can follow; slower)")
    ap.add_argument("--plain-fold", action="store_true", help="GEMM-folded form without the epilogue-fused layout (FoldedNet)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak: every GPU plays --games games; strong: --total-games games are divided over the GPUs")
    ap.add_argument("--total-games", type=int, default=0, help="with --scaling strong: games of the whole job (default --games)")
    ap.add_argument("--preset", choices=["config4"], default=None,
                    help="config4 = BASELINE.json configs[3]: 131,072 concurrent games x 1,600 sims/move, 64-filter ResNet, bf16")
    ap.add_argument("--nn", choices=["native", "cublas"], default="native",
                    help="bf16 network: native = the library's tcgen05 kernel (csrc/net.cu), cublas = the same folded "
                         "network as PyTorch/cuBLASLt GEMMs in bucketed CUDA graphs")
    args = ap.parse_args()
    if args.preset == "config4":
        args.games, args.sims, args.width = 131072, 1600, 64
    return args


def make_model(width: int, dtype: torch.dtype, device):
    from c4a0_b200.nn import ConnectFourNet, ModelConfig

    torch.manual_seed(1337)  # tests/c4a0_tests/conftest.py:8 of the reference
    cfg = ModelConfig(n_residual_blocks=1, conv_filter_size=width, n_policy_layers=4, n_value_layers=2)
    return ConnectFourNet(cfg).to(device=device, dtype=dtype).eval()


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu: int):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "250", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


# ------------------------------------------------------------------------------------------------
# CPU reference arm: oracle/selfplay_threads.cpp (the reference's thread/channel architecture)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(n_games: int, n_iter: int, width: int, nn_device: str, max_sims: int = 0):
    """The oracle's threaded port of rust/src/self_play.rs on the host cores.  max_sims > 0 stops
    the run after that many simulations (time-boxed sample at full concurrency).
    Returns dict(positions, moves, sims, seconds, threads, nn_batches)."""
    import oracle

    model = make_model(width, torch.float32, nn_device)
    L = oracle.lib()
    bits = np.arange(42, dtype=np.uint64)[None, :]

    def cb(_user, model_id, n, pos, policy, qp, qn):  # rust/src/pybridge.rs:170-198
        arr = np.ctypeslib.as_array(C.cast(pos, C.POINTER(C.c_uint64)), shape=(n, 2))
        mask, value = arr[:, 0], arr[:, 1]
        mine = ((mask & value)[:, None] >> bits) & np.uint64(1)
        theirs = ((mask & ~value)[:, None] >> bits) & np.uint64(1)
        planes = np.concatenate([mine, theirs], axis=1).astype(np.float32).reshape(n, 2, 6, 7)
        pol, a, b = model.forward_numpy(planes)
        C.memmove(policy, pol.ctypes.data, 28 * n)
        C.memmove(qp, a.ctypes.data, 4 * n)
        C.memmove(qn, b.ctypes.data, 4 * n)

    fn = oracle.EVAL_FN(cb)
    md = (oracle.Metadata * n_games)(*[oracle.Metadata(i, 0, 0) for i in range(n_games)])
    out = (oracle.Sample * (n_games * oracle.MAX_SAMPLES))()
    out_n = (C.c_int * n_games)()
    st = oracle.Stats()
    nb = C.c_uint64(0)
    threads = max(1, (os.cpu_count() or 2) - 1)  # self_play.rs:78
    t0 = time.perf_counter()
    rc = L.c4o_self_play_threaded_budget(
        md, n_games, 2000, n_iter, C_EXPLORATION, C_PLY_PENALTY, C.cast(fn, C.c_void_p), None, threads, max_sims,
        out, out_n, C.byref(st), C.byref(nb),
    )
    dt = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"CPU reference self-play failed rc={rc}")
    finished = sum(1 for i in range(n_games) if out_n[i] > 0)
    return dict(positions=int(st.samples), moves=int(st.moves), finished=finished, sims=int(st.sims), seconds=dt,
                threads=threads + 1, nn_batches=int(nb.value))


SPEC_ROOM = 8192  # rows of max_nn_batch_size above the resident games: used for speculative evaluations
CPU_GAMES = 1000  # BASELINE.json configs[0]: the reference's own CPU-runnable case (1,000 games x 600 sims)


def _nn_device(args) -> str:
    return "cuda:0" if torch.cuda.is_available() and getattr(args, "cpu_nn_device", "auto") != "cpu" else "cpu"


def cpu_sample(args, budget_s: float, rate_hint: float = 0.0) -> dict:
    """Time-boxed sample of configs[0] at its full concurrency (1,000 games in flight, which sets the
    reference's NN batch size): the first `budget` simulations of the job.  Returns the raw run."""
    nn_device = _nn_device(args)
    if rate_hint <= 0.0:
        cpu_reference_run(CPU_GAMES, args.sims, args.width, nn_device, max_sims=100_000)  # warm-up: threads, CUDA context
        probe = cpu_reference_run(CPU_GAMES, args.sims, args.width, nn_device, max_sims=400_000)
        rate_hint = probe["sims"] / probe["seconds"]
    budget = int(max(300_000, rate_hint * budget_s))
    return cpu_reference_run(CPU_GAMES, args.sims, args.width, nn_device, max_sims=budget)


def cpu_whole_job(args) -> dict:
    """configs[0] played to completion: every game from the empty board to its terminal position."""
    run = cpu_reference_run(CPU_GAMES, args.sims, args.width, _nn_device(args), max_sims=0)
    run["positions_per_s"] = run["positions"] / run["seconds"]
    run["sims_per_s"] = run["sims"] / run["seconds"]
    run["sims_per_position"] = run["sims"] / max(1, run["positions"])
    return run


def _cpu_sample_text(args, what: str) -> str:
    return (f"{what} of {CPU_GAMES} concurrent games x {args.sims} sims/move (BASELINE configs[0]); threaded port of "
            f"rust/src/self_play.rs on the host cores, fp32 network via the numpy callback on {_nn_device(args)}")


def cpu_baseline(args, budget_s: float) -> dict:
    """The `cpu_baseline` object of our own bench line: a time-boxed sample (the first simulations of
    configs[0]).  Early moves cost more simulations per position than a whole game does (no subtree to
    reuse yet), so positions/s is quoted as sims/s divided by the WHOLE-JOB simulations per position of
    the GPU run of the same workload shape (passed in by the caller as args._sims_per_position) when
    known; the raw early-game count is kept next to it."""
    run = cpu_sample(args, budget_s)
    sims_per_s = run["sims"] / run["seconds"]
    early = (run["moves"] + run["finished"]) / run["seconds"]
    spp = getattr(args, "_sims_per_position", None)
    return {
        "value": sims_per_s / spp if spp else early,
        "unit": UNIT,
        "cores": run["threads"],
        "kind": "port",
        "sample": _cpu_sample_text(args, f"first {run['sims']} simulations ({run['seconds']:.1f} s)")
                  + ("; positions/s = sims/s / whole-job simulations per position of this workload "
                     f"({spp:.1f}, from the GPU run: identical games)" if spp else "; positions = moves made + games finished"),
        "sims_per_s": sims_per_s,
        "early_game_positions_per_s": early,
        "seconds": run["seconds"],
        "host_cpus": os.cpu_count(),
        "mean_nn_batch": (run["sims"] / run["nn_batches"]) if run["nn_batches"] else None,
    }


def run_reference(args):
    """`--impl reference`: the reference's CPU self-play architecture on the host cores.  Timed step 1 plays
    configs[0] to completion (all 1,000 games, ~30 s), so positions/s covers whole games; the remaining
    steps are time-boxed samples of the same job (its first simulations) whose sims/s is converted with the
    complete job's simulations per position.  `value` = positions/s over the whole timed region."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total_budget = 170.0
    for _ in range(args.warmup):
        cpu_sample(args, 1.0, rate_hint=300_000.0)
    t0 = time.perf_counter()
    full = cpu_whole_job(args)
    rest = max(0, args.steps - 1)
    per_step = max(2.0, min(20.0, (total_budget - full["seconds"]) / max(1, rest)))
    sims, secs = full["sims"], full["seconds"]
    rates = [full["sims_per_s"]]
    for _ in range(rest):
        r = cpu_sample(args, per_step, rate_hint=full["sims_per_s"])
        sims += r["sims"]
        secs += r["seconds"]
        rates.append(r["sims"] / r["seconds"])
    dt = time.perf_counter() - t0
    sims_per_s = sims / secs
    v = sims_per_s / full["sims_per_position"]
    cfg = workload_config(args, 1)
    cfg.update({
        "workload": f"gen-0 self-play, {CPU_GAMES} concurrent games x {args.sims} MCTS sims/move (BASELINE configs[0], the "
                    f"reference's CPU-runnable case), 7x6 board, random-init c4a0 ResNet (1 block x {args.width} filters, "
                    f"4 policy / 2 value layers), c_exploration={C_EXPLORATION}, c_ply_penalty={C_PLY_PENALTY}",
        "games_per_gpu": CPU_GAMES, "global_games": CPU_GAMES, "games": CPU_GAMES,
        "time_boxed": f"step 1 of {args.steps} plays all {CPU_GAMES} games to completion; the others are the first "
                      f"~{per_step:.0f} s of the same job",
        "parallelism": f"{full['threads']} host threads (1 NN thread + workers), network on {_nn_device(args)}",
        "l2": "n/a (CPU arm)",
    })
    base = {
        "value": v, "unit": UNIT, "cores": full["threads"], "kind": "port",
        "sample": _cpu_sample_text(args, f"one complete job ({full['sims']} simulations, {full['positions']} positions, "
                                         f"{full['seconds']:.1f} s) + {rest} time-boxed samples"),
        "sims_per_s": sims_per_s, "host_cpus": os.cpu_count(),
        "whole_job": {"games": CPU_GAMES, "finished": full["finished"], "positions": full["positions"], "sims": full["sims"],
                      "seconds": full["seconds"], "positions_per_s": full["positions_per_s"], "sims_per_s": full["sims_per_s"],
                      "sims_per_position": full["sims_per_position"]},
        "step_sims_per_s": rates,
        "mean_nn_batch": (full["sims"] / full["nn_batches"]) if full["nn_batches"] else None,
    }
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sims_per_s": sims_per_s,
    }
    print(json.dumps(line), flush=True)


def latest_counters():
    """profiles/rNN_kernel_counters.json of the latest round (tools/ncu_counters.py over `ncu --set full` captures of
    the benched job's own launches)."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9]_kernel_counters.json")))
    if not files:
        return None, None
    try:
        return json.load(open(files[-1])), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


def workload_config(args, world):
    games = getattr(args, "_games_per_gpu", args.games)
    return {
        "workload": f"gen-0 self-play, {games} lockstep games per GPU x {args.sims} MCTS sims/move, 7x6 board, "
                    f"random-init c4a0 ResNet (1 block x {args.width} filters, 4 policy / 2 value layers), "
                    f"c_exploration={C_EXPLORATION}, c_ply_penalty={C_PLY_PENALTY}",
        "games_per_gpu": games, "sims_per_move": args.sims, "global_games": games * world,
        "max_nn_batch_size": games + SPEC_ROOM,
        "parallelism": f"games sharded over {world} GPU(s), no search-path collective",
        "l2": "inputs larger than L2: the games' live trees (tens of GB of arenas per GPU, ~1.5 GB of them touched "
              "between two visits of a game) and the evaluation cache (GBs; see engine_device_gb) against 126 MB of L2; "
              "every step starts with an empty evaluation cache",
    }


def nn_roofline(args, flops_per_row, rows_per_tick, nn_ms, native):
    """The network kernel against the tensor roofline: reference-form FLOPs of the rows a tick really evaluates over
    the kernel's sampled device time, against the measured dense bf16 peak; tensor-pipe utilisation by ncu counter
    from the round's capture of the benched job."""
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        src = "measured burst (MEASURED_PEAKS.json)"
    except Exception:
        peak, src = 2250.0, "nominal dense bf16 (B200_PROFILING.md)"
    out = {"bound": "tensor", "unit": "TFLOP/s", "flops_per_eval": flops_per_row, "peak": peak, "peak_source": src,
           "kernel": "k_net2 (csrc/net.cu: tcgen05.mma cta_group::2, TMEM accumulators, TMA operand ring)" if native
                     else "cuBLASLt GEMM chain via PyTorch (library)",
           "rows_per_launch": rows_per_tick, "avg_launch_ms": nn_ms}
    if nn_ms:
        out["achieved"] = flops_per_row * rows_per_tick / (nn_ms * 1e-3) / 1e12
        out["frac"] = out["achieved"] / peak
    counters, counters_file = latest_counters()
    if native and counters and "k_net2" in counters:
        kn = counters["k_net2"]
        pct = kn.get("tensor_pipe_active_pct_of_elapsed") or []
        if pct:
            out["tensor_pipe_active_pct"] = sum(pct) / len(pct)  # sm__pipe_tensor_cycles_active, % of elapsed cycles
            out["tensor_pipe_source"] = counters_file
            out["l2_to_sm_bytes_per_launch"] = (sum(kn.get("l2_to_sm_read") or [0]) / max(1, len(kn.get("l2_to_sm_read") or [1])))
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from c4a0_b200 import dist as D
    from c4a0_b200 import selfplay
    from c4a0_b200.selfplay import DeviceEvaluator
    import c4a0_rust

    rank, world, local_rank = D.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dtype = torch.bfloat16 if args.nn_dtype == "bf16" else torch.float32
    model = make_model(args.width, torch.float32, device)
    selfplay.DEFAULTS["sample_kernels_every"] = args.sample_kernels_every
    selfplay.DEFAULTS["n_lanes"] = args.lanes
    selfplay.DEFAULTS["host_loop"] = args.host_loop
    selfplay.DEFAULTS["dedup"] = not args.no_dedup
    selfplay.DEFAULTS["max_inline_sims"] = args.max_inline
    selfplay.DEFAULTS["eval_cache"] = not args.no_eval_cache
    if args.arena_mult > 0:
        selfplay.DEFAULTS["arena_blocks"] = int(args.arena_mult * (args.sims + 2))
    selfplay.DEFAULTS["eval_cache_entries"] = args.eval_cache_entries
    selfplay.DEFAULTS["speculate"] = not args.no_speculate
    selfplay.DEFAULTS["spec_rows"] = args.spec_rows
    if args.scaling == "strong":  # a fixed job divided over the GPUs (contiguous game-id ranges, like dist.shard_range)
        total = args.total_games or args.games
        lo, hi = D.shard_range(total, rank, world)
        G = hi - lo
        ids = range(lo, hi)
        args._games_per_gpu = (total + world - 1) // world
    else:  # weak scaling: every rank plays its own G games
        G = args.games
        ids = range(rank * G, (rank + 1) * G)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    state = {"ev": None}

    def one_step():
        t0 = time.perf_counter()
        D.broadcast_model(model)  # a generation's weights: rank 0 -> all (NCCL); no-op at N=1
        # weights -> inference form, loaded in place into the previous generation's evaluator so the
        # captured CUDA graphs and the engine of the previous call are reused
        evaluator = DeviceEvaluator.from_model(model, dtype, fold=(False if args.no_fold else ("plain" if args.plain_fold else (True if args.nn == "native" else "cublas"))), reuse=state["ev"])
        state["ev"] = evaluator
        reqs = [c4a0_rust.GameMetadata(i, 0, 0) for i in ids]
        # max_nn_batch_size: the resident games plus room for the speculative rows (play_games keeps
        # n_slots + spec_rows within the caller's bound)
        # the trainer lives on rank 0: at N > 1 every rank keeps its samples on the device; the valid ones travel
        # packed, straight from the engines' sample stores, to rank 0 (NCCL send/recv), which copies them to the host
        selfplay.DEFAULTS["fetch"] = world == 1  # at N > 1 the gather below delivers rank 0's own games as well
        with torch.cuda.nvtx.range("play_games"):
            res = c4a0_rust.play_games(reqs, G + SPEC_ROOM, args.sims, C_EXPLORATION, C_PLY_PENALTY, evaluator)
        n_pos = int(res._run_info.stats["samples"])
        checksum = float(res._soa.q_no_penalty.sum()) if world == 1 else 0.0  # touch the host result
        if world > 1:
            meta = np.array([(i, 0, 0) for i in ids], dtype=np.uint64)
            gm, gc, gp = D.gather_session_samples(c4a0_rust._native._SESSION["sess"], meta, packed_result=True)
            if rank == 0:  # host arrays: per game its id and sample count, per valid sample 52 bytes
                state["gathered"] = int(gc.sum())
                n_all = (args.total_games or args.games) if args.scaling == "strong" else G * world
                assert state["gathered"] >= n_pos and len(gm) == n_all and gp.shape == (state["gathered"], D.PACK_WORDS)
                checksum += float(gp[:, 12].view(np.float32).sum())
        return time.perf_counter() - t0, res._run_info, n_pos, checksum

    first_call_s = None
    for i in range(args.warmup):
        t_first = time.perf_counter()
        one_step()
        if i == 0:  # engine creation (arenas, cache, network buffers), first launches: what the session cache hides later
            torch.cuda.synchronize()
            first_call_s = time.perf_counter() - t_first
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    t_start = time.perf_counter()
    runs = [one_step() for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t_start
    clk = clocks.stop() if rank == 0 else None

    dev_s = sum(r[1].device_s for r in runs)
    e2e_s = sum(r[0] for r in runs)
    positions = sum(r[2] for r in runs)
    sims = sum(r[1].stats["sims"] for r in runs)
    evals = sum(r[1].stats["nn_evals"] for r in runs)
    expansions = sum(r[1].stats["expansions"] for r in runs)
    # rows the network kernels were launched on: bucket sizes for PyTorch graphs; the library's kernel reads the
    # tick's own row count on the device, so it runs on exactly the rows evaluated
    nn_rows_launched = (sum(r[1].stats["nn_evals"] for r in runs) if type(state["ev"]).__name__ == "NativeEvaluator"
                        else sum(r[1].report.get("nn_rows_launched", 0) for r in runs))
    compactions = sum(r[1].stats.get("compactions", 0) for r in runs)
    cache_hits = sum(r[1].stats.get("cache_hits", 0) for r in runs)
    cache_inserts = sum(r[1].stats.get("cache_inserts", 0) for r in runs)
    spec_rows = sum(r[1].stats.get("spec_rows", 0) for r in runs)
    engine_gb = runs[-1][1].engine_bytes / 1e9
    ticks = sum(r[1].ticks for r in runs)
    depth = sum(r[1].stats["select_depth_sum"] for r in runs)
    kms = [r[1].kernel_ms for r in runs if r[1].kernel_ms]
    # max over ranks of the times, sum over ranks of the work
    t = torch.tensor([dev_s, e2e_s, wall], dtype=torch.float64, device=device)
    w = torch.tensor([positions, sims, evals, expansions, nn_rows_launched], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(w, op=torch.distributed.ReduceOp.SUM)
    dev_s_max, e2e_s_max, wall_max = t.tolist()
    positions_all, sims_all, evals_all, expansions_all, rows_launched_all = w.tolist()
    if rank != 0:
        return

    # roofline of the tree kernel (k_step): SURVEY.md §8(d) bytes per simulation
    d = depth / max(1, sims)
    e = expansions / max(1, sims)
    s = 2 if dtype == torch.bfloat16 else 4
    bytes_per_sim = 92 * d + 24 * (d + 1) + 16 + 84 * s + 36 + 168 * e
    sims_per_launch = sims / max(1, ticks)
    peak, peak_src = measured_peak_gbs()
    roofline = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                "peak_source": peak_src, "kernel": "k_step (network answer -> softmax/expand/backup, move, UCT select, dedup + plane pack, tick close)"}
    if kms:
        k_step_ms = float(np.mean([k["k_step"] for k in kms]))
        achieved = bytes_per_sim * sims_per_launch / (k_step_ms * 1e-3) / 1e9
        roofline.update(
            achieved=achieved, frac=achieved / peak, bytes_per_sim=bytes_per_sim, sims_per_launch=sims_per_launch,
            avg_launch_ms=k_step_ms, nn_graph_avg_ms=float(np.mean([k.get("nn", 0.0) for k in kms])), launches_timed=int(sum(k["samples"] for k in kms)),
            select_depth=d, expand_frac=e,
        )
        counters, counters_file = latest_counters()
        if counters and "k_step" in counters:
            # dram__bytes_read.sum + dram__bytes_write.sum of k_step launches of THIS job (a mid-job tick of the default
            # bench configuration) under `ncu --set full`; cold-cache per-launch figure, see profiles/
            ks = counters["k_step"]
            roofline["traffic"] = ks.get("dram_bytes_per_launch")
            roofline["traffic_source"] = counters_file
            if ks.get("sims_in_captured_launch"):
                roofline["traffic_over_algorithmic"] = ks["dram_bytes_per_launch"] / (bytes_per_sim * ks["sims_in_captured_launch"])
    flops = model.flops_per_position()
    native = type(state["ev"]).__name__ == "NativeEvaluator"
    nn_form = ("module" if args.no_fold else "GEMM-folded (FoldedNet)" if args.plain_fold else
               "one persistent tcgen05/TMEM/TMA kernel over the GEMM-folded network (k_net2, csrc/net.cu)" if native else
               "GEMM-folded, epilogue-fused (FusedNet), cuBLASLt")
    line = {
        "metric": METRIC, "value": positions_all / dev_s_max, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dev_s_max / max(1, args.steps), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": args.nn_dtype, "tree_dtype": "f32 (bit-exact with the reference), u64 bitboards",
        "data": "synthetic", "config": workload_config(args, world),
        "sims_per_s": sims_all / dev_s_max, "nn_evals_per_s": evals_all / dev_s_max,
        "e2e": {
            "value": positions_all / e2e_s_max, "unit": UNIT,
            "h2d_bytes_per_step": 3 * 8 * G,
            # N = 1: the games as padded arrays; N > 1 (rank 0): every rank's valid samples packed at 52 B each
            "d2h_bytes_per_step": (G * (4 + 43 * (8 + 8 + 28 + 4 + 4)) if world == 1
                                   else int(52 * positions_all / max(1, args.steps)) + 28 * G * world),
            "sims_per_s": sims_all / e2e_s_max, "api": "c4a0_rust.play_games(list[GameMetadata], ...) -> PlayGamesResult",
            "wall_ms_per_step": 1e3 * wall_max / max(1, args.steps),
            # the very first call of the process (untimed warm-up step 1): device allocations of the engine, the
            # evaluation cache and the network, first kernel launches; later calls reuse the cached session
            "first_call_ms": None if first_call_s is None else 1e3 * first_call_s,
        },
        # our kernels inside the timed region: k_step once per tick, k_tail where a tick needed it,
        # k_init_globals + k_init + k_tail per call, k_sum_counters per call, and k_head_epilogue (the
        # network's output stage) once per network graph launch when the folded forms are used
        "gpu_launches": int(ticks + sum(r[1].report.get("tail_launches", 0) for r in runs) + 4 * args.steps * args.lanes
                            + (0 if args.no_fold else sum(r[1].report.get("nn_launches", 0) for r in runs))),
        "leaf_evals_per_s": expansions_all / dev_s_max,
        "dedup": {"enabled": not args.no_dedup, "leaf_requests": expansions_all, "unique_rows": evals_all,
                  "rows_launched_incl_bucket_padding": rows_launched_all},
        # rank 0's counters; the table is emptied at the start of every step (= every play_games call)
        "eval_cache": {"enabled": not args.no_eval_cache, "hits": cache_hits, "inserts": cache_inserts,
                       "hit_rate_of_expansions": cache_hits / max(1, expansions),
                       "speculate": not (args.no_eval_cache or args.no_speculate), "speculative_rows": spec_rows},
        "compactions_per_step": compactions / max(1, args.steps), "engine_device_gb": engine_gb,
        "lanes": args.lanes, "nn_form": nn_form,
        "roofline": roofline,
        "nn_roofline": nn_roofline(args, flops, evals / max(1, ticks), roofline.get("nn_graph_avg_ms"), native),
        "host_loop": {"wait_ms_per_step": sum(r[1].report.get("host_wait_ms", 0.0) for r in runs) / max(1, args.steps),
                      "launch_ms_per_step": sum(r[1].report.get("host_launch_ms", 0.0) for r in runs) / max(1, args.steps),
                      "tail_launches_per_step": sum(r[1].report.get("tail_launches", 0) for r in runs) / max(1, args.steps),
                      "nn_relaunches_per_step": sum(r[1].report.get("nn_relaunches", 0) for r in runs) / max(1, args.steps)},
        "clocks": clk,
        "bucket_launches": [int(sum(r[1].report.get("bucket_launches", [0] * 32)[i] for r in runs)) for i in range(20)],
        "ticks_per_step": ticks / max(1, args.steps),
    }
    if world == 1 and not args.no_eval_cache and not args.no_ablation:
        # the same job with every leaf sent to the network (untimed warm-up step first: new engine, new graphs)
        try:
            selfplay.DEFAULTS["eval_cache"] = False
            one_step()
            torch.cuda.synchronize()
            dt, info, n_pos, _ = one_step()
            line["without_eval_cache"] = {
                "value": n_pos / info.device_s, "unit": UNIT, "ms_per_step": 1e3 * info.device_s, "steps": 1,
                "e2e_value": n_pos / dt, "ticks_per_step": info.ticks, "nn_rows": info.stats["nn_evals"],
            }
        except Exception as exc:
            line["without_eval_cache"] = {"value": None, "error": str(exc)}
        finally:
            selfplay.DEFAULTS["eval_cache"] = True
            c4a0_rust._native.close_cached_session()
    if world == 1 and args.nn_dtype == "bf16" and not args.no_ablation:
        # the same job with the network evaluated in f32 (the reference arm's precision)
        try:
            c4a0_rust._native.close_cached_session()
            state["ev"], dtype = None, torch.float32
            one_step()
            torch.cuda.synchronize()
            dt, info, n_pos, _ = one_step()
            line["f32_network"] = {
                "value": n_pos / info.device_s, "unit": UNIT, "ms_per_step": 1e3 * info.device_s, "steps": 1,
                "e2e_value": n_pos / dt, "ticks_per_step": info.ticks, "sims_per_s": info.stats["sims"] / info.device_s,
            }
        except Exception as exc:
            line["f32_network"] = {"value": None, "error": str(exc)}
        finally:
            dtype = torch.bfloat16
            state["ev"] = None
            c4a0_rust._native.close_cached_session()
    if world == 1 and not args.no_cpu_baseline:
        try:
            args._sims_per_position = sims / max(1, positions)
            line["cpu_baseline"] = cpu_baseline(args, args.cpu_seconds)
        except Exception as exc:  # the number is a report, never a reason to lose the bench line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {exc}"}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
