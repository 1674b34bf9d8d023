"""The GEMM-folded inference form computes the same function as the module form (CPU, fp32/fp64)."""

import numpy as np
import pytest
import torch

from c4a0_b200.nn import ConnectFourNet, FoldedNet, FusedNet, ModelConfig, default_config


def _randomize_bn(model):
    g = torch.Generator().manual_seed(5)
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.3)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)


def _planes(n, seed=0):
    rng = np.random.default_rng(seed)
    cells = rng.integers(0, 3, size=(n, 42))
    x = np.zeros((n, 2, 42), np.float32)
    x[:, 0][cells == 1] = 1
    x[:, 1][cells == 2] = 1
    return torch.from_numpy(x.reshape(n, 2, 6, 7))


@pytest.mark.parametrize(
    "cfg",
    [
        dict(n_residual_blocks=1, conv_filter_size=8, n_policy_layers=4, n_value_layers=2),
        dict(n_residual_blocks=3, conv_filter_size=4, n_policy_layers=2, n_value_layers=3),
        dict(n_residual_blocks=0, conv_filter_size=4, n_policy_layers=1, n_value_layers=1),
        dict(n_residual_blocks=2, conv_filter_size=4, n_policy_layers=1, n_value_layers=2),
    ],
)
def test_folded_equals_module(cfg):
    torch.manual_seed(11)
    model = ConnectFourNet(ModelConfig(**cfg)).double().eval()
    _randomize_bn(model)
    x = _planes(64).double()
    with torch.no_grad():
        want = model(x)
        got = FoldedNet(model, dtype=torch.float64)(x)
    for a, b in zip(got, want):
        assert torch.allclose(a.double(), b, atol=1e-9), (a - b).abs().max()


def test_default_net_shapes_and_flops():
    torch.manual_seed(1337)
    model = ConnectFourNet(default_config()).eval()
    assert model.fc_size == 1344
    assert model.flops_per_position() == 16_071_552
    assert sum(p.numel() for p in model.parameters()) > 7_000_000
    f = FoldedNet(model, dtype=torch.float32)
    x = _planes(8)
    with torch.no_grad():
        a, b = model(x), f(x)
    for u, v in zip(a, b):
        assert torch.allclose(u, v, atol=2e-4)
    pol = b[0].exp().sum(1)
    assert torch.allclose(pol, torch.ones_like(pol), atol=1e-5)
    assert b[1].abs().max() <= 1 and b[2].abs().max() <= 1


def test_state_dict_names_match_reference_layout():
    keys = set(ConnectFourNet(default_config()).state_dict().keys())
    for k in (
        "conv.0.weight", "conv.1.block.0.weight", "conv.1.block.1.bias", "conv.1.block.2.running_mean",
        "fc_policy.0.0.weight", "fc_policy.0.1.running_var", "fc_policy.3.weight", "fc_value.0.0.bias", "fc_value.1.weight",
    ):
        assert k in keys, k


def test_fused_form_equals_module_and_refreshes_in_place():
    torch.manual_seed(21)
    cfg = dict(n_residual_blocks=1, conv_filter_size=6, n_policy_layers=4, n_value_layers=2)
    model = ConnectFourNet(ModelConfig(**cfg)).double().eval()
    _randomize_bn(model)
    x = _planes(32).double()
    f = FusedNet(model, dtype=torch.float64)
    assert (f.plane_stride, f.plane_offset) == (model.fc_size + 96, model.fc_size)
    buf = torch.zeros(32, f.plane_stride, dtype=torch.float64)
    buf[:, f.plane_offset : f.plane_offset + 84] = x.reshape(32, 84)
    with torch.no_grad():
        want = model(x)
        got = f(buf)
        got4 = f(x)
    for a, b, c in zip(got, want, got4):
        assert torch.allclose(a.double(), b, atol=1e-6) and torch.allclose(c.double(), b, atol=1e-6)
    # a new generation of weights, loaded in place
    torch.manual_seed(22)
    model2 = ConnectFourNet(ModelConfig(**cfg)).double().eval()
    ptr = f.f_w2.data_ptr()
    f.refresh(model2)
    assert f.f_w2.data_ptr() == ptr
    buf[:, : f.plane_offset] = 0
    with torch.no_grad():
        for a, b in zip(f(buf), model2(x)):
            assert torch.allclose(a.double(), b, atol=1e-6)
    assert not FusedNet.supports(ConnectFourNet(ModelConfig(n_residual_blocks=2, conv_filter_size=2, n_policy_layers=2, n_value_layers=2)))


def test_folding_leaves_the_callers_module_alone():
    """handing a module that is being trained to the engine must not flip it to eval mode.  The fold
    reads parameters and running statistics only, so it is the same in either mode."""
    from c4a0_b200.nn import ConnectFourNet, FoldedNet, ModelConfig

    torch.manual_seed(5)
    m = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=4, n_policy_layers=2, n_value_layers=2))
    for mod in m.modules():
        if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 1.5)
    m.train()
    a, meta_a = FoldedNet._fold(m)
    assert m.training and all(mod.training for mod in m.modules())
    m.eval()
    b, meta_b = FoldedNet._fold(m)
    assert meta_a == meta_b and all(torch.equal(a[k], b[k]) for k in a)
