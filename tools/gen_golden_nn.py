"""Generates tests/golden/nn_reference_forward.npz by importing the REFERENCE's network class.

Runs only in the build container (needs /root/reference).  pytorch_lightning / torchmetrics are not
installed here, so minimal stand-ins are registered before the import: `pl.LightningModule` ->
`torch.nn.Module` (+ no-op `save_hyperparameters`), torchmetrics classes -> parameter-free modules.
Nothing of the reference's forward pass is replaced: conv stack, heads, BatchNorm, LogSoftmax and
Tanh are the reference's own code (src/c4a0/nn.py:59-117, 184-195).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)  # our c4a0_rust provides N_COLS / N_ROWS for the reference's import

pl = types.ModuleType("pytorch_lightning")


class LightningModule(torch.nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    @property
    def device(self):  # Lightning's property
        return next(self.parameters()).device


pl.LightningModule = LightningModule
tm = types.ModuleType("torchmetrics")


class _Metric(torch.nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


tm.KLDivergence = _Metric
tm.MeanSquaredError = _Metric
sys.modules["pytorch_lightning"] = pl
sys.modules["torchmetrics"] = tm
import c4a0_rust  # noqa: E402,F401  ours, imported first so the reference's stub package is not picked up

sys.path.insert(0, "/root/reference/src")
from c4a0.nn import ConnectFourNet, ModelConfig  # noqa: E402  (the reference's class)

out = {}
for tag, cfg in {
    "a": dict(n_residual_blocks=1, conv_filter_size=4, n_policy_layers=4, n_value_layers=2),
    "b": dict(n_residual_blocks=2, conv_filter_size=3, n_policy_layers=2, n_value_layers=3),
}.items():
    torch.manual_seed(1337)
    model = ConnectFourNet(ModelConfig(lr_schedule={0: 1e-3}, l2_reg=1e-4, **cfg))
    g = torch.Generator().manual_seed(99)
    for m in model.modules():  # non-trivial BatchNorm statistics, as after training
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.2)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    rng = np.random.default_rng(7)
    cells = rng.integers(0, 3, size=(16, 42))
    x = np.zeros((16, 2, 42), np.float32)
    x[:, 0][cells == 1] = 1
    x[:, 1][cells == 2] = 1
    x = x.reshape(16, 2, 6, 7)
    pol, qp, qn = model.forward_numpy(x)
    out[f"{tag}_cfg"] = np.array([cfg["n_residual_blocks"], cfg["conv_filter_size"], cfg["n_policy_layers"], cfg["n_value_layers"]])
    out[f"{tag}_x"], out[f"{tag}_policy"], out[f"{tag}_qp"], out[f"{tag}_qn"] = x, pol, qp, qn
    for k, v in model.state_dict().items():
        out[f"{tag}_sd/{k}"] = v.numpy()
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "nn_reference_forward.npz"), **out)
print("wrote", len(out), "arrays")
