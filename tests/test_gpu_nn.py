"""GPU tests of the network's fused output layers / output stage (c4a0_heads, c4a0_head_epilogue;
reference nn.py:100-130)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


@pytest.mark.parametrize("form,dtype", [("folded", "float32"), ("fused", "float32"), ("fused", "bfloat16")])
@pytest.mark.parametrize("rows", [1, 127, 4096, 5000])
def test_fused_output_stage_matches_torch(form, dtype, rows):
    _need_gpu()
    from c4a0_b200.nn import ConnectFourNet, FoldedNet, FusedNet, default_config

    torch.manual_seed(7)
    dt = getattr(torch, dtype)
    model = ConnectFourNet(default_config()).cuda().eval()
    net = (FusedNet if form == "fused" else FoldedNet)(model, dtype=dt)
    stride = net.plane_stride if form == "fused" else FoldedNet.IN_PAD
    off = net.plane_offset if form == "fused" else 0
    buf = torch.zeros(rows, stride, device="cuda", dtype=dt)
    buf[:, off : off + 84] = (torch.rand(rows, 84, device="cuda") < 0.3).to(dt)
    with torch.no_grad():
        pol, a, b = net(buf.clone())
        logits = torch.full((rows, 7), 9.0, device="cuda")
        qp = torch.full((rows,), 9.0, device="cuda")
        qn = torch.full((rows,), 9.0, device="cuda")
        out = net(buf.clone(), out=(logits, qp, qn))
    assert out[0] is logits
    torch.cuda.synchronize()
    # the fused kernel keeps the output layers in f32; PyTorch rounds them to the activations' dtype first
    tol = 2e-2 if dtype == "bfloat16" else 2e-5
    np.testing.assert_allclose(logits.cpu().numpy(), pol.float().cpu().numpy(), atol=tol, rtol=0)
    np.testing.assert_allclose(qp.cpu().numpy(), a.float().cpu().numpy(), atol=tol, rtol=0)
    np.testing.assert_allclose(qn.cpu().numpy(), b.float().cpu().numpy(), atol=tol, rtol=0)
    # and against the module itself in f32
    with torch.no_grad():
        mp, ma, mb = model(buf[:, off : off + 84].float().reshape(rows, 2, 6, 7))
    mtol = 6e-2 if dtype == "bfloat16" else 1e-4
    np.testing.assert_allclose(logits.cpu().numpy(), mp.cpu().numpy(), atol=mtol, rtol=0)
    np.testing.assert_allclose(qp.cpu().numpy(), ma.cpu().numpy(), atol=mtol, rtol=0)
    np.testing.assert_allclose(qn.cpu().numpy(), mb.cpu().numpy(), atol=mtol, rtol=0)
    assert np.allclose(np.exp(logits.cpu().numpy()).sum(1), 1.0, atol=1e-5)


def test_head_epilogue_rejects_bad_buffers():
    _need_gpu()
    from c4a0_b200.nn import _output_stage

    pol = torch.zeros(4, 8, device="cuda")
    val = torch.zeros(4, 8, device="cuda")
    good = (torch.zeros(4, 7, device="cuda"), torch.zeros(4, device="cuda"), torch.zeros(4, device="cuda"))
    _output_stage(pol, val, out=good)
    with pytest.raises(ValueError):
        _output_stage(pol, val, out=(torch.zeros(3, 7, device="cuda"), good[1], good[2]))
    with pytest.raises(ValueError):
        _output_stage(pol, val, out=(good[0].double(), good[1], good[2]))
