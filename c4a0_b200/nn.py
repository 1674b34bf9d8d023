"""The c4a0 policy/value network as a plain torch module (no Lightning).

The network stays PyTorch by design (BASELINE.json north star): the engine only hands it a
device tensor of input planes and reads its outputs in place.  This file restates the reference
architecture (src/c4a0/nn.py:59-117, 184-195) with the SAME submodule names — `conv`,
`conv.N.block`, `fc_policy`, `fc_value` — so `state_dict()`s interchange with the reference's
`ConnectFourNet` (whose extra torchmetrics members hold no parameters).  pytorch_lightning and
torchmetrics are not installed in this image, hence the restatement.

Outputs follow the reference: `policy` are log-probabilities (LogSoftmax), `q_penalty` and
`q_no_penalty` are tanh values; MCTS re-normalises the policy over legal moves itself
(c4r.rs:272-286 + mcts.rs:416-434), so any logits-like tensor works.
"""

from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
from pydantic import BaseModel
from torch import nn

N_ROWS, N_COLS = 6, 7


class ModelConfig(BaseModel):
    """Same fields as the reference's ModelConfig (src/c4a0/nn.py:16-38)."""

    n_residual_blocks: int
    conv_filter_size: int
    n_policy_layers: int
    n_value_layers: int
    lr_schedule: Dict[int, float] = {0: 2e-3}
    l2_reg: float = 4e-4


def default_config() -> ModelConfig:
    """CLI defaults of `main.py train` (src/c4a0/main.py:46-56)."""
    return ModelConfig(n_residual_blocks=1, conv_filter_size=32, n_policy_layers=4, n_value_layers=2)


class ResidualBlock(nn.Module):
    """x + relu(bn(conv(conv(x)))) — src/c4a0/nn.py:184-195."""

    def __init__(self, n_channels: int, kernel_size: int = 3, padding: int = 1) -> None:
        super().__init__()
        self.block = nn.Sequential(
            nn.Conv2d(n_channels, n_channels, kernel_size=kernel_size, padding=padding),
            nn.Conv2d(n_channels, n_channels, kernel_size=kernel_size, padding=padding),
            nn.BatchNorm2d(n_channels),
            nn.ReLU(),
        )

    def forward(self, x):
        return x + self.block(x)


def _head(fc_size: int, n_layers: int, n_out: int, final: nn.Module) -> nn.Sequential:
    hidden = [
        nn.Sequential(nn.Linear(fc_size, fc_size), nn.BatchNorm1d(fc_size), nn.ReLU()) for _ in range(n_layers - 1)
    ]
    return nn.Sequential(*hidden, nn.Linear(fc_size, n_out), final)


class ConnectFourNet(nn.Module):
    EPS = 1e-8

    def __init__(self, config: ModelConfig):
        super().__init__()
        self.config = config
        c = config.conv_filter_size
        self.conv = nn.Sequential(
            nn.Conv2d(2, c, kernel_size=3, padding=1),
            *[ResidualBlock(c) for _ in range(config.n_residual_blocks)],
        )
        fc_size = c * N_ROWS * N_COLS
        self.fc_size = fc_size
        self.fc_policy = _head(fc_size, config.n_policy_layers, N_COLS, nn.LogSoftmax(dim=1))
        self.fc_value = _head(fc_size, config.n_value_layers, 2, nn.Tanh())

    def forward(self, x) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        x = self.conv(x).flatten(1)  # b (c h w)
        policy_logprobs = self.fc_policy(x)
        q = self.fc_value(x)
        return policy_logprobs, q[:, 0], q[:, 1]

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def forward_numpy(self, x: np.ndarray):
        """numpy in / numpy out, eval mode — the callback the reference's trainer passes to
        play_games (src/c4a0/nn.py:119-130, training.py:188)."""
        self.eval()
        p0 = next(self.parameters())
        with torch.no_grad():
            pol, a, b = self.forward(torch.from_numpy(x).to(device=p0.device, dtype=p0.dtype))
        return tuple(np.ascontiguousarray(t.float().cpu().numpy()) for t in (pol, a, b))

    def flops_per_position(self) -> int:
        """2*MACs of one forward pass (convs + linears), the figure BASELINE.md §3 quotes."""
        total = 0
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                total += 2 * m.in_channels * m.out_channels * m.kernel_size[0] * m.kernel_size[1] * N_ROWS * N_COLS
            elif isinstance(m, nn.Linear):
                total += 2 * m.in_features * m.out_features
        return total
