"""The layer program c4a0_b200.native_net folds a ConnectFourNet into (what the sm_100a kernel k_net runs)
is the same function as the module (reference: src/c4a0/nn.py:59-117 in eval mode) — checked in float64 on
the CPU by executing the program's semantics in plain torch."""

import pytest
import torch

from c4a0_b200 import _lib as L
from c4a0_b200.native_net import emulate_program, fold_program
from c4a0_b200.nn import ConnectFourNet, ModelConfig


def _model(width, n_p, n_v, seed=0):
    torch.manual_seed(seed)
    m = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=width, n_policy_layers=n_p, n_value_layers=n_v)).double()
    # non-trivial BatchNorm statistics, as after training
    for mod in m.modules():
        if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.2)
    return m.eval()


@pytest.mark.parametrize("width,n_p,n_v", [(8, 2, 2), (8, 4, 2), (16, 3, 3), (32, 4, 2)])
def test_program_equals_module(width, n_p, n_v):
    m = _model(width, n_p, n_v)
    layers, n_buffers = fold_program(m)
    planes = (torch.rand(37, 2, 6, 7, dtype=torch.float64) < 0.3).double()
    want = m(planes)
    got = emulate_program(layers, n_buffers, planes.reshape(37, 84))
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=1e-9, rtol=1e-9)


def test_program_shapes_obey_the_kernel_contract():
    layers, n_buffers = fold_program(_model(32, 4, 2))
    names = [l[0] for l in layers]
    assert names == ["l1", "l2", "p0", "vf", "p1", "pf"] and n_buffers == 4
    Fp = 1344
    for i, (name, W, b, m) in enumerate(layers):
        assert W.shape[1] % L.NET_TILE_K == 0 and b.shape[0] == W.shape[0]
        assert W.shape[0] == (L.NET_HEAD_N if m["kind"] != L.NET_HIDDEN else W.shape[0] // L.NET_PAD_N * L.NET_PAD_N)
        assert -1 <= m["dep"] < i and m["inp"][1] % 64 == 0
    assert layers[1][1].shape == (2 * Fp, Fp + 128)
    # width 8: F = 336 is padded to 1344 columns
    layers, _ = fold_program(_model(8, 2, 2))
    assert layers[0][1].shape == (1344, 128) and layers[1][1].shape == (2688, 1472)
