"""Generation training loop around the GPU self-play engine — Lightning-free.

The reference alternates self-play and supervised training (src/c4a0/training.py:155-294) with
PyTorch Lightning, which is not installed in this image.  This module restates that
loop with plain torch, keeps the on-disk layout of a generation
(`<base>/<isoformat timestamp>/{metadata.json, games.pkl, model.pkl}`, training.py:41-67, 95-134) and
calls self-play through the drop-in boundary `c4a0_rust.play_games` — on the device fast path.

Training semantics mirrored from the reference:
  * loss = KL(policy_target || policy) + MSE(q_penalty) + MSE(q_no_penalty), with
    log(target + 1e-8) as the policy target in log space (src/c4a0/nn.py:160-172);
  * Adam(lr = lr_schedule[gen], weight_decay = l2_reg) (nn.py:140-152);
  * whole-game 80/20 split with seed 1337, horizontal-flip augmentation of both sets
    (training.py:207, 316-317);
  * at most 100 epochs, early stopping on val_loss with patience 10, the best epoch's weights are
    kept (training.py:210-222, utils.py:35-93).
Under torchrun every rank plays a contiguous shard of the games; samples are gathered to rank 0,
which trains; the new weights are broadcast over NCCL.
"""

from __future__ import annotations

import copy
import os
import pickle
from datetime import datetime
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from pydantic import BaseModel

import c4a0_rust
from c4a0_rust import PlayGamesResult

from . import dist as D
from .nn import ConnectFourNet, ModelConfig


class TrainingGen(BaseModel):
    """One generation's metadata (same fields as the reference's TrainingGen, training.py:25-39)."""

    created_at: datetime
    gen_n: int
    n_mcts_iterations: int
    c_exploration: float
    c_ply_penalty: float
    self_play_batch_size: int
    training_batch_size: int
    parent: Optional[datetime] = None
    val_loss: Optional[float] = None
    solver_score: Optional[float] = None

    def gen_folder(self, base_dir: str) -> str:
        return os.path.join(base_dir, self.created_at.isoformat())

    def save_all(self, base_dir: str, games: Optional[PlayGamesResult], model: ConnectFourNet) -> None:
        d = self.gen_folder(base_dir)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "metadata.json"), "w") as f:
            f.write(self.model_dump_json(indent=2))
        with open(os.path.join(d, "games.pkl"), "wb") as f:
            pickle.dump(games, f)
        cpu_model = copy.deepcopy(model).cpu()
        with open(os.path.join(d, "model.pkl"), "wb") as f:
            pickle.dump(cpu_model, f)
        # the same weights without a class path: what a reference checkout (c4a0.nn.ConnectFourNet, a
        # LightningModule) can load with ConnectFourNet(ModelConfig(**config)).load_state_dict(state_dict)
        torch.save({"config": cpu_model.config.model_dump(), "state_dict": cpu_model.state_dict()},
                   os.path.join(d, "model_state.pt"))

    @staticmethod
    def load_all(base_dir: str) -> List["TrainingGen"]:
        if not os.path.isdir(base_dir):
            return []
        stamps = []
        for name in os.listdir(base_dir):
            if os.path.isdir(os.path.join(base_dir, name)):
                try:
                    stamps.append(datetime.fromisoformat(name))
                except ValueError:
                    continue
        out = []
        for t in sorted(stamps, reverse=True):
            with open(os.path.join(base_dir, t.isoformat(), "metadata.json")) as f:
                out.append(TrainingGen.model_validate_json(f.read()))
        return out

    @staticmethod
    def load_latest_with_default(base_dir: str, model_config: ModelConfig, **params) -> "TrainingGen":
        gens = TrainingGen.load_all(base_dir)
        if gens:
            return gens[0]
        gen = TrainingGen(created_at=datetime.now(), gen_n=0, **params)
        gen.save_all(base_dir, None, ConnectFourNet(model_config))  # gen 0 = random init
        return gen

    def get_model(self, base_dir: str) -> ConnectFourNet:
        """model.pkl of this generation.  A directory written by the reference holds a pickled
        `c4a0.nn.ConnectFourNet` (a LightningModule with torchmetrics members, training.py:62-67); its
        submodules `conv`, `fc_policy`, `fc_value` are plain torch and carry the same parameter names, so the
        weights are moved into this package's class (see load_reference_model)."""
        d = self.gen_folder(base_dir)
        model = None
        if os.path.exists(os.path.join(d, "model.pkl")):
            with open(os.path.join(d, "model.pkl"), "rb") as f:
                try:
                    model = _ModelUnpickler(f).load()
                except Exception:
                    model = None
        if isinstance(model, ConnectFourNet) and "_c4a0_foreign" not in model.__dict__:
            return model
        if model is not None:
            return load_reference_model(model)
        state = os.path.join(d, "model_state.pt")
        if os.path.exists(state):
            blob = torch.load(state, map_location="cpu", weights_only=False)
            out = ConnectFourNet(ModelConfig(**blob["config"]))
            out.load_state_dict(blob["state_dict"])
            return out.eval()
        raise RuntimeError(f"cannot load {os.path.join(d, 'model.pkl')}")

    def get_games(self, base_dir: str) -> Optional[PlayGamesResult]:
        with open(os.path.join(self.gen_folder(base_dir), "games.pkl"), "rb") as f:
            return pickle.load(f)


class _Stub:
    """Stands in for classes of packages this image lacks (pytorch_lightning, torchmetrics, ...) while
    unpickling a model the reference wrote: accepts any constructor arguments and any state."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)

    def __call__(self, *a, **k):
        return _Stub()


class _ModelUnpickler(pickle.Unpickler):
    """Maps the reference's class paths onto this package's classes; unknown third-party classes become stubs."""

    def find_class(self, module, name):
        if module in ("c4a0.nn", "src.c4a0.nn"):
            from . import nn as local

            if name == "ConnectFourNet":
                return _ForeignNet
            if hasattr(local, name):
                return getattr(local, name)
            return _Stub
        try:
            return super().find_class(module, name)
        except (ImportError, AttributeError):
            return _Stub


class _ForeignNet(torch.nn.Module):
    """Receives the pickled state of the reference's ConnectFourNet (whatever its base classes stored)."""

    def __setstate__(self, state):
        torch.nn.Module.__init__(self)
        self.__dict__.update(state)
        self.__dict__["_c4a0_foreign"] = True


def load_reference_model(obj) -> ConnectFourNet:
    """A model unpickled from a reference-written model.pkl -> this package's ConnectFourNet with the same
    weights.  The hyper-parameters come from the object's `config` (ModelConfig fields, nn.py:16-38) or are
    read off the layer shapes."""
    mods = obj.__dict__.get("_modules", {})
    conv, fc_policy, fc_value = mods.get("conv"), mods.get("fc_policy"), mods.get("fc_value")
    if conv is None or fc_policy is None or fc_value is None:
        raise RuntimeError("the pickled model has no conv / fc_policy / fc_value submodules")
    cfg = obj.__dict__.get("config")
    if isinstance(cfg, dict):
        cfg = ModelConfig(**cfg)
    if not isinstance(cfg, ModelConfig):
        fields = {k: getattr(cfg, k) for k in ("n_residual_blocks", "conv_filter_size", "n_policy_layers", "n_value_layers")
                  if cfg is not None and hasattr(cfg, k)}
        if len(fields) != 4:  # read the architecture off the modules
            fields = dict(n_residual_blocks=len(list(conv.children())) - 1, conv_filter_size=list(conv.children())[0].out_channels,
                          n_policy_layers=len(list(fc_policy.children())) - 1, n_value_layers=len(list(fc_value.children())) - 1)
        extra = {k: getattr(cfg, k) for k in ("lr_schedule", "l2_reg") if cfg is not None and hasattr(cfg, k)}
        cfg = ModelConfig(**fields, **extra)
    out = ConnectFourNet(cfg)
    sd = {}
    for prefix, m in (("conv", conv), ("fc_policy", fc_policy), ("fc_value", fc_value)):
        for k, v in m.state_dict().items():
            sd[f"{prefix}.{k}"] = v
    out.load_state_dict(sd)
    return out.eval()


def parse_lr_schedule(floats: List[float]) -> Dict[int, float]:
    """[0, 2e-3, 10, 8e-4] -> {0: 2e-3, 10: 8e-4} (training.py:349-360)."""
    if len(floats) % 2:
        raise ValueError("lr_schedule must have an even number of elements")
    out = {}
    for g, lr in zip(floats[::2], floats[1::2]):
        if int(g) != g:
            raise ValueError("lr_schedule alternates generation (int) and learning rate")
        out[int(g)] = float(lr)
    return out


def lr_for_generation(schedule: Dict[int, float], gen_n: int) -> float:
    """The rate of the last threshold <= gen_n (nn.py:140-152)."""
    items = sorted(schedule.items())
    lr = items[0][1]
    for threshold, rate in items[1:]:
        if gen_n < threshold:
            break
        lr = rate
    return lr


def loss_terms(model: ConnectFourNet, pos, policy_target, qp_target, qn_target):
    """(total, policy KL, q_penalty MSE, q_no_penalty MSE) — nn.py:160-172."""
    logp, qp, qn = model(pos)
    log_target = torch.log(policy_target + ConnectFourNet.EPS)
    kl = torch.sum(log_target.exp() * (log_target - logp), dim=-1).mean()
    mse_p = torch.mean((qp - qp_target) ** 2)
    mse_n = torch.mean((qn - qn_target) ** 2)
    return kl + mse_p + mse_n, kl, mse_p, mse_n


Arrays = Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]


def split_arrays(games: PlayGamesResult, train_frac: float = 0.8, seed: int = 1337, augment: bool = True) -> Tuple[Arrays, Arrays]:
    """PlayGamesResult.split_train_test (whole games, pybridge.rs:107-120) + flip augmentation
    (training.py:316-317), on arrays instead of one Python object per sample."""
    from .engine import GameSamples, host_shuffle

    soa, meta = games._soa, games._meta
    n = len(soa.n_samples)
    order = host_shuffle(seed, n)
    x = np.float32(n) * np.float32(train_frac)
    n_train = int(max(0, min(n, np.floor(np.abs(x) + np.float32(0.5)))))

    def subset(idx) -> Arrays:
        sub = PlayGamesResult._from_soa(
            meta[idx],
            GameSamples(*[getattr(soa, f)[idx] for f in ("n_samples", "mask", "value", "policy", "q_penalty", "q_no_penalty")]),
        )
        pos, pol, qp, qn = sub.to_arrays()
        if augment:  # mirror image: columns reversed in both planes and in the policy
            pos = np.concatenate([pos, pos[:, :, :, ::-1]])
            pol = np.concatenate([pol, pol[:, ::-1]])
            qp, qn = np.concatenate([qp, qp]), np.concatenate([qn, qn])
        return np.ascontiguousarray(pos), np.ascontiguousarray(pol), qp, qn

    return subset(order[:n_train]), subset(order[n_train:])


def fit(
    model: ConnectFourNet,
    train: Arrays,
    val: Arrays,
    batch_size: int,
    lr: float,
    l2_reg: float,
    device: torch.device,
    max_epochs: int = 100,
    patience: int = 10,
    seed: int = 1337,
    log=None,
) -> Tuple[ConnectFourNet, float, int]:
    """Returns (best model, its val_loss, epochs run)."""
    model = copy.deepcopy(model).to(device=device, dtype=torch.float32)
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=l2_reg)
    tr = [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in train]
    va = [torch.from_numpy(np.ascontiguousarray(a)).to(device) for a in val]
    n_tr, n_va = len(tr[0]), len(va[0])
    if n_tr == 0 or n_va == 0:
        raise ValueError("training and validation sets must not be empty")
    g = torch.Generator(device="cpu").manual_seed(seed)
    best_state, best_loss, bad_epochs, epochs = None, float("inf"), 0, 0
    for epoch in range(max_epochs):
        epochs = epoch + 1
        model.train()  # BatchNorm statistics are trained too (training.py:221)
        perm = torch.randperm(n_tr, generator=g).to(device)
        for lo in range(0, n_tr, batch_size):
            idx = perm[lo : lo + batch_size]
            if len(idx) < 2:
                continue  # BatchNorm1d needs more than one value per channel in training mode
            loss, *_ = loss_terms(model, *[t[idx] for t in tr])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        model.eval()
        total = torch.zeros((), dtype=torch.float64, device=device)
        with torch.no_grad():
            for lo in range(0, n_va, batch_size):
                part = [t[lo : lo + batch_size] for t in va]
                total += loss_terms(model, *part)[0].double() * len(part[0])
        val_loss = float(total) / n_va  # one host sync per epoch
        if log:
            log(f"epoch {epoch}: val_loss {val_loss:.5f}")
        if val_loss < best_loss:
            best_loss, bad_epochs = val_loss, 0
            best_state = {k: v.clone() for k, v in model.state_dict().items()}
        else:
            bad_epochs += 1
            if bad_epochs >= patience:
                break
    model.load_state_dict(best_state)
    return model.eval(), best_loss, epochs


def self_play(model: ConnectFourNet, n_games: int, batch_size: int, n_mcts_iterations: int, c_exploration: float,
              c_ply_penalty: float, device: torch.device, nn_dtype: torch.dtype = torch.bfloat16,
              evaluator=None):
    """This rank's shard of the generation's games through c4a0_rust.play_games (device fast path).
    Returns (PlayGamesResult of ALL games on rank 0 / None elsewhere, evaluator for reuse)."""
    from .selfplay import DeviceEvaluator

    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    model = model.to(device=device, dtype=torch.float32)
    D.broadcast_model(model)
    evaluator = DeviceEvaluator.from_model(model, nn_dtype, reuse=evaluator)
    lo, hi = D.shard_range(n_games, rank, world)
    reqs = [c4a0_rust.GameMetadata(i, 0, 0) for i in range(lo, hi)]  # training.py:181
    from . import selfplay

    if world == 1:
        return c4a0_rust.play_games(reqs, batch_size, n_mcts_iterations, c_exploration, c_ply_penalty, evaluator), evaluator
    # every rank keeps its samples on the device; the valid ones travel packed to rank 0 (dist.py)
    selfplay.DEFAULTS["fetch"] = False
    try:
        c4a0_rust.play_games(reqs, batch_size, n_mcts_iterations, c_exploration, c_ply_penalty, evaluator)
    finally:
        selfplay.DEFAULTS["fetch"] = True
    meta = np.array([(i, 0, 0) for i in range(lo, hi)], dtype=np.uint64).reshape(-1, 3)
    meta, soa = D.gather_session_samples(c4a0_rust._native._SESSION["sess"], meta)
    games = PlayGamesResult._from_soa(meta, soa) if rank == 0 else None
    return games, evaluator


def train_single_gen(base_dir: str, device: torch.device, parent: TrainingGen, n_self_play_games: int,
                     n_mcts_iterations: int, c_exploration: float, c_ply_penalty: float, self_play_batch_size: int,
                     training_batch_size: int, model_config: Optional[ModelConfig] = None, max_epochs: int = 100,
                     nn_dtype: torch.dtype = torch.bfloat16, evaluator=None, log=print, report: Optional[list] = None):
    """One generation: self-play with the parent's model, train a copy on the samples, save.  `report` (rank 0)
    collects one dict per generation: wall seconds of self-play (all ranks, gather included), training and saving."""
    import time

    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    gen_n = parent.gen_n + 1
    model = parent.get_model(base_dir)
    cfg = model_config or model.config
    t0 = time.perf_counter()
    torch.cuda.nvtx.range_push(f"c4a0.gen{gen_n}.self_play")
    games, evaluator = self_play(model, n_self_play_games, self_play_batch_size, n_mcts_iterations, c_exploration,
                                 c_ply_penalty, device, nn_dtype, evaluator)
    torch.cuda.nvtx.range_pop()
    t1 = time.perf_counter()
    gen = None
    if rank == 0:
        torch.cuda.nvtx.range_push(f"c4a0.gen{gen_n}.train")
        train, val = split_arrays(games, 0.8, 1337, augment=True)
        lr = lr_for_generation(cfg.lr_schedule, gen_n)
        if log:
            log(f"gen {gen_n}: {int(games._soa.n_samples.sum())} samples, {games.unique_positions()} unique positions, lr {lr}")
        best, val_loss, epochs = fit(model, train, val, training_batch_size, lr, cfg.l2_reg, device, max_epochs=max_epochs, log=None)
        gen = TrainingGen(
            created_at=datetime.now(), gen_n=gen_n, n_mcts_iterations=n_mcts_iterations, c_exploration=c_exploration,
            c_ply_penalty=c_ply_penalty, self_play_batch_size=self_play_batch_size, training_batch_size=training_batch_size,
            parent=parent.created_at, val_loss=val_loss,
        )
        t2 = time.perf_counter()
        gen.save_all(base_dir, games, best)
        t3 = time.perf_counter()
        torch.cuda.nvtx.range_pop()
        n_pos = int(games._soa.n_samples.sum())
        if log:
            log(f"gen {gen_n}: val_loss {val_loss:.5f} after {epochs} epochs | self-play {t1 - t0:.2f} s "
                f"({n_pos / (t1 - t0):.0f} positions/s on {world} GPU(s)), training {t2 - t1:.2f} s, save {t3 - t2:.2f} s")
        if report is not None:
            report.append(dict(gen=gen_n, games=n_self_play_games, sims_per_move=n_mcts_iterations, gpus=world, positions=n_pos,
                               self_play_s=t1 - t0, positions_per_s=n_pos / (t1 - t0), train_s=t2 - t1, epochs=epochs,
                               save_s=t3 - t2, val_loss=val_loss))
    if torch.distributed.is_initialized():
        box = [gen]
        torch.distributed.broadcast_object_list(box, src=0)
        gen = box[0]
        torch.distributed.barrier()
    return gen, evaluator


def training_loop(base_dir: str, device: torch.device, n_self_play_games: int, n_mcts_iterations: int,
                  c_exploration: float, c_ply_penalty: float, self_play_batch_size: int, training_batch_size: int,
                  model_config: ModelConfig, max_gens: Optional[int] = None, max_epochs: int = 100,
                  nn_dtype: torch.dtype = torch.bfloat16, log=print, report: Optional[list] = None) -> TrainingGen:
    """training.py:242-294: generation after generation until max_gens."""
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    params = dict(n_mcts_iterations=n_mcts_iterations, c_exploration=c_exploration, c_ply_penalty=c_ply_penalty,
                  self_play_batch_size=self_play_batch_size, training_batch_size=training_batch_size)
    if rank == 0:
        gen = TrainingGen.load_latest_with_default(base_dir, model_config, **params)
    else:
        gen = None
    if torch.distributed.is_initialized():
        box = [gen]
        torch.distributed.broadcast_object_list(box, src=0)
        gen = box[0]
        torch.distributed.barrier()
    evaluator = None
    while True:  # like the reference (training.py:278-294): at least one generation, then check max_gens
        gen, evaluator = train_single_gen(
            base_dir, device, gen, n_self_play_games, model_config=model_config, max_epochs=max_epochs,
            nn_dtype=nn_dtype, evaluator=evaluator, log=log if rank == 0 else None, report=report, **params,
        )
        if max_gens is not None and gen.gen_n >= max_gens:
            return gen


def main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(description="c4a0 training loop on the B200 self-play engine (main.py train)")
    ap.add_argument("--base-dir", default="training")
    ap.add_argument("--n-self-play-games", type=int, default=1700)  # main.py:40-49 defaults
    ap.add_argument("--n-mcts-iterations", type=int, default=1400)
    ap.add_argument("--c-exploration", type=float, default=6.6)
    ap.add_argument("--c-ply-penalty", type=float, default=0.01)
    ap.add_argument("--self-play-batch-size", type=int, default=2000)
    ap.add_argument("--training-batch-size", type=int, default=2000)
    ap.add_argument("--n-residual-blocks", type=int, default=1)
    ap.add_argument("--conv-filter-size", type=int, default=32)
    ap.add_argument("--n-policy-layers", type=int, default=4)
    ap.add_argument("--n-value-layers", type=int, default=2)
    ap.add_argument("--lr-schedule", type=float, nargs="+", default=[0, 2e-3, 10, 8e-4])
    ap.add_argument("--l2-reg", type=float, default=4e-4)
    ap.add_argument("--max-gens", type=int, default=None)
    ap.add_argument("--max-epochs", type=int, default=100)
    ap.add_argument("--nn-dtype", choices=["bf16", "f32"], default="bf16")
    ap.add_argument("--report", default=None, help="write one JSON line per generation (timings) to this file")
    ap.add_argument("--tf32", action="store_true",
                    help="train with TF32 tensor-core matmuls / convolutions (torch.set_float32_matmul_precision('high')). "
                         "Off by default: the reference trains in plain fp32, and a training step is 0.17 TFLOP of fp32 GEMMs")
    a = ap.parse_args(argv)
    rank, world, local_rank = D.init_from_env()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if a.tf32:
        torch.set_float32_matmul_precision("high")
        torch.backends.cudnn.allow_tf32 = True
    cfg = ModelConfig(n_residual_blocks=a.n_residual_blocks, conv_filter_size=a.conv_filter_size,
                      n_policy_layers=a.n_policy_layers, n_value_layers=a.n_value_layers,
                      lr_schedule=parse_lr_schedule(a.lr_schedule), l2_reg=a.l2_reg)
    import time

    report: list = []
    t0 = time.perf_counter()
    gen = training_loop(a.base_dir, device, a.n_self_play_games, a.n_mcts_iterations, a.c_exploration, a.c_ply_penalty,
                        a.self_play_batch_size, a.training_batch_size, cfg, a.max_gens, a.max_epochs,
                        torch.bfloat16 if a.nn_dtype == "bf16" else torch.float32, report=report)
    if rank == 0:
        print(f"finished at generation {gen.gen_n}, val_loss {gen.val_loss}, {time.perf_counter() - t0:.1f} s wall for "
              f"{len(report)} generation(s) on {world} GPU(s)")
        if a.report:
            import json

            with open(a.report, "w") as f:
                for r in report:
                    f.write(json.dumps(r) + "\n")
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
