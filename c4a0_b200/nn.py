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


# ------------------------------------------------------------------------------------------------
# Inference form: the same function as a short chain of GEMMs
# ------------------------------------------------------------------------------------------------
def _conv_as_matrix(conv: nn.Conv2d):
    """A 3x3/pad-1 convolution on the fixed 6x7 board is a linear map R^{Cin*42} -> R^{Cout*42}:
    returns (M [Cin*42, Cout*42], bias [Cout*42]) in float64, exact (zero padding included)."""
    cin, cout = conv.in_channels, conv.out_channels
    w = conv.weight.detach().double()
    eye = torch.eye(cin * 42, dtype=torch.float64, device=w.device).reshape(cin * 42, cin, N_ROWS, N_COLS)
    m = torch.nn.functional.conv2d(eye, w, None, conv.stride, conv.padding).reshape(cin * 42, cout * 42)
    b = conv.bias.detach().double() if conv.bias is not None else torch.zeros(cout, dtype=torch.float64, device=w.device)
    return m, b.repeat_interleave(42)


def _bn_scale_shift(bn, repeat: int = 1):
    """Eval-mode BatchNorm as y = x * s + t (float64)."""
    s = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    t = bn.bias.detach().double() - bn.running_mean.detach().double() * s
    return s.repeat_interleave(repeat), t.repeat_interleave(repeat)


def _output_stage(pol: torch.Tensor, val: torch.Tensor, out=None):
    """log_softmax over the 7 policy outputs, tanh of the two values (reference nn.py:116-130) from the
    last linear layers' outputs `pol` [B, >=7] and `val` [B, >=2].  With out = (logits [B,7] f32,
    q_penalty [B] f32, q_no_penalty [B] f32) on CUDA, one kernel of the engine library writes them in
    place (c4a0_head_epilogue) and `out` is returned."""
    if out is None:
        logits = pol[:, :7].float()
        q = torch.tanh(val[:, :2].float())
        return torch.log_softmax(logits, dim=1), q[:, 0], q[:, 1]
    from . import _lib as L

    logits, qp, qn = out
    rows = pol.shape[0]
    ok = (pol.is_cuda and pol.dtype == val.dtype and pol.dtype in (torch.float32, torch.bfloat16)
          and pol.stride(1) == 1 and val.stride(1) == 1
          and all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape[0] == rows for t in out))
    if not ok:
        raise ValueError("output buffers must be contiguous float32 CUDA tensors with one row per input row")
    L.check(L.lib().c4a0_head_epilogue(
        pol.data_ptr(), val.data_ptr(), L.PLANES_BF16 if pol.dtype == torch.bfloat16 else L.PLANES_F32,
        pol.stride(0), val.stride(0), rows, logits.data_ptr(), qp.data_ptr(), qn.data_ptr(),
        torch.cuda.current_stream(pol.device).cuda_stream))
    return out


class FoldedNet(nn.Module):
    """Inference-only restatement of a trained/initialised `ConnectFourNet` as dense GEMMs.

    Same function, same weights, evaluated in an order that suits tensor cores:
      * every convolution acts on a fixed 6x7 board, so it is a constant matrix; eval-mode
        BatchNorm is an affine map and folds into the layer before it;
      * the reference's residual block is x + relu(bn(conv(conv(x)))) with NO nonlinearity between
        its two convolutions (src/c4a0/nn.py:184-195), so the pair is one matrix; for the first
        block, whose input is itself affine in the 84 input planes, block-input and
        block-preactivation come out of ONE [B,84(+pad)] x [84, 2*F] GEMM;
      * the first hidden layers of the policy and value heads read the same activations and are
        concatenated into one GEMM.
    For the default net (1 block x 32 filters) a forward pass is 7 GEMMs, 4 of them 1344-wide.
    Outputs match the module form to rounding (tests/test_nn_fold.py); PyTorch (cuBLASLt) stays
    the runtime — this class only reorganises the weights at load time.
    `in_features` is the padded input width (a multiple of 8 so bf16 rows are 16-byte aligned);
    the engine writes its planes with that row stride.
    """

    IN_PAD = 96
    FUSED_HEADS_MAX_ROWS = 4096

    def __init__(self, model: ConnectFourNet, dtype: torch.dtype = torch.bfloat16, device=None):
        super().__init__()
        device = device if device is not None else next(model.parameters()).device
        self.F, self.dtype = model.fc_size, dtype
        tensors, meta = self._fold(model)
        self.n_blocks, self.joint_first, self.n_p, self.n_v = meta
        for name, t in tensors.items():
            self.register_buffer(name, t.to(device=device, dtype=dtype).contiguous(), persistent=False)
        # biases of the output layers in f32 for the fused output kernel
        self.register_buffer("bpf32", tensors["bpf"].to(device=device, dtype=torch.float32).contiguous(), persistent=False)
        self.register_buffer("bvf32", tensors["bvf"].to(device=device, dtype=torch.float32).contiguous(), persistent=False)

    @torch.no_grad()
    def refresh(self, model: ConnectFourNet) -> "FoldedNet":
        """Load a new generation's weights IN PLACE (same architecture): CUDA graphs captured over
        this module stay valid, so nothing is re-captured between generations."""
        tensors, meta = self._fold(model)
        if meta != (self.n_blocks, self.joint_first, self.n_p, self.n_v) or model.fc_size != self.F:
            raise ValueError("refresh() needs a model of the same architecture")
        for name, t in tensors.items():
            getattr(self, name).copy_(t)
        self.bpf32.copy_(tensors["bpf"])
        self.bvf32.copy_(tensors["bvf"])
        return self

    @classmethod
    @torch.no_grad()
    def _fold(cls, model: ConnectFourNet):
        """float64 folded weights {buffer name: tensor} + structure (n_blocks, joint_first, n_p, n_v).  The fold is
        the eval-mode function (BatchNorm with its running statistics); only parameters and buffers are read, the
        module's training flag is left as found."""
        F = model.fc_size
        out = {}
        layers = list(model.conv.children())
        stem, blocks = layers[0], layers[1:]
        w_in, b_in = _conv_as_matrix(stem)  # x = planes @ w_in + b_in   [84 -> F]

        def block_affine(blk):
            ca, cb, bn, act = list(blk.block.children())
            assert isinstance(ca, nn.Conv2d) and isinstance(cb, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d)
            assert isinstance(act, nn.ReLU)
            ma, ba = _conv_as_matrix(ca)
            mb, bb = _conv_as_matrix(cb)
            s, t = _bn_scale_shift(bn, 42)
            return (ma @ mb) * s[None, :], (ba @ mb + bb) * s + t

        n_blocks = len(blocks)
        if blocks:
            wb, bb_ = block_affine(blocks[0])
            w0 = torch.cat([w_in @ wb, w_in], dim=1)  # [84, 2F]: block pre-activation | block input
            b0 = torch.cat([b_in @ wb + bb_, b_in])
        else:
            w0, b0 = w_in, b_in
        w0p = torch.zeros(cls.IN_PAD, w0.shape[1], dtype=torch.float64, device=w0.device)
        w0p[:84] = w0
        out["w0"], out["b0"] = w0p, b0
        for j, blk in enumerate(blocks[1:], start=1):
            out[f"wblk{j}"], out[f"bblk{j}"] = block_affine(blk)

        def head_layers(seq):
            hidden, final = [], None
            for m in seq.children():
                if isinstance(m, nn.Sequential):
                    lin, bn, act = list(m.children())
                    s, t = _bn_scale_shift(bn)
                    w = lin.weight.detach().double().t() * s[None, :]
                    b = lin.bias.detach().double() * s + t
                    hidden.append((w, b))
                elif isinstance(m, nn.Linear):
                    final = (m.weight.detach().double().t(), m.bias.detach().double())
            return hidden, final

        ph, pf = head_layers(model.fc_policy)
        vh, vf = head_layers(model.fc_value)
        joint_first = bool(ph) and bool(vh)
        if joint_first:
            out["wh0"] = torch.cat([ph[0][0], vh[0][0]], dim=1)
            out["bh0"] = torch.cat([ph[0][1], vh[0][1]])
            ph, vh = ph[1:], vh[1:]
        for i, (w, b) in enumerate(ph):
            out[f"wp{i}"], out[f"bp{i}"] = w, b
        for i, (w, b) in enumerate(vh):
            out[f"wv{i}"], out[f"bv{i}"] = w, b
        # final layers, output width padded to 8 columns
        dev0 = pf[0].device
        wpf = torch.zeros(F, 8, dtype=torch.float64, device=dev0)
        wpf[:, :7] = pf[0]
        bpf = torch.zeros(8, dtype=torch.float64, device=dev0)
        bpf[:7] = pf[1]
        wvf = torch.zeros(F, 8, dtype=torch.float64, device=dev0)
        wvf[:, :2] = vf[0]
        bvf = torch.zeros(8, dtype=torch.float64, device=dev0)
        bvf[:2] = vf[1]
        out.update(wpf=wpf, bpf=bpf, wvf=wvf, bvf=bvf)
        # the same two layers transposed, for the fused output kernel (c4a0_heads): [7][F], [2][F]
        out["wpf_t"] = pf[0].t().contiguous()
        out["wvf_t"] = vf[0].t().contiguous()
        return out, (n_blocks, joint_first, len(ph), len(vh))

    @staticmethod
    def _lin_relu(x, w, b):
        if x.is_cuda:
            return torch._addmm_activation(b, x, w)  # cuBLASLt bias+ReLU epilogue
        return torch.relu(torch.addmm(b, x, w))

    def forward(self, planes: torch.Tensor, out=None):
        """planes: [B, IN_PAD] (engine layout, zero padded) or [B,2,6,7]."""
        if planes.dim() == 4:
            x0 = planes.new_zeros(planes.shape[0], self.IN_PAD)
            x0[:, :84] = planes.reshape(planes.shape[0], 84)
        else:
            x0 = planes
        F = self.F
        z = torch.addmm(self.b0, x0, self.w0)
        h = torch.relu(z[:, :F]) + z[:, F:] if self.n_blocks else z
        for j in range(1, self.n_blocks):
            h = self._lin_relu(h, getattr(self, f"wblk{j}"), getattr(self, f"bblk{j}")) + h
        if self.joint_first:
            y = self._lin_relu(h, self.wh0, self.bh0)
            hp, hv = y[:, :F], y[:, F:]
        else:
            hp = hv = h
        for i in range(self.n_p):
            hp = self._lin_relu(hp, getattr(self, f"wp{i}"), getattr(self, f"bp{i}"))
        for i in range(self.n_v):
            hv = self._lin_relu(hv, getattr(self, f"wv{i}"), getattr(self, f"bv{i}"))
        return self._heads(hp, hv, out)

    def _heads(self, hp: torch.Tensor, hv: torch.Tensor, out=None):
        """Output layers + output stage from the heads' last hidden activations.  With `out` (the engine's
        logits / q buffers) on CUDA this is one kernel of the engine library (c4a0_heads); otherwise two
        GEMMs and the PyTorch output stage."""
        # measured on B200 (tools/nn_curve.py): the CUDA-core kernel wins below ~4,096 rows (one launch
        # instead of three: 28.7 vs 32.8 us at 64 rows, 43 vs 49 us at 1,024), the tensor-core GEMMs above
        if out is not None and hp.is_cuda and hp.shape[0] <= self.FUSED_HEADS_MAX_ROWS:
            from . import _lib as L

            logits, qp, qn = out
            rows = hp.shape[0]
            fits = (hp.dtype == hv.dtype == self.dtype and self.dtype in (torch.float32, torch.bfloat16)
                    and hp.stride(1) == 1 and hv.stride(1) == 1 and self.F % 8 == 0
                    and hp.stride(0) % 8 == 0 and hv.stride(0) % 8 == 0
                    and hp.data_ptr() % 16 == 0 and hv.data_ptr() % 16 == 0
                    and all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape[0] == rows for t in out))
            if fits:
                L.check(L.lib().c4a0_heads(
                    hp.data_ptr(), hv.data_ptr(), L.PLANES_BF16 if self.dtype == torch.bfloat16 else L.PLANES_F32,
                    hp.stride(0), hv.stride(0), self.F, self.wpf_t.data_ptr(), self.bpf32.data_ptr(),
                    self.wvf_t.data_ptr(), self.bvf32.data_ptr(), rows, logits.data_ptr(), qp.data_ptr(), qn.data_ptr(),
                    torch.cuda.current_stream(hp.device).cuda_stream))
                return out
        return _output_stage(torch.addmm(self.bpf, hp, self.wpf), torch.addmm(self.bvf, hv, self.wvf), out)

    def flops_per_position(self) -> int:
        total = 0
        for name, t in self.named_buffers():
            if name.startswith("w") and not name.endswith("_t"):
                total += 2 * t.shape[0] * t.shape[1]
        return total


class FusedNet(FoldedNet):
    """FoldedNet with every elementwise pass folded into a GEMM epilogue (default c4a0 family:
    one residual block, both heads with at least one hidden layer).

    The block output is h = relu(pre) + inp with `pre` and `inp` both affine in the 84 input planes.
    The layer after it computes h @ W = relu(pre) @ W + inp @ W, and inp @ W is again affine in the
    planes: inp @ W = planes @ (W_in W) + b_in W.  So with the activation buffer laid out as
    [ relu(pre) (F columns) | planes (96 columns) ] per row,

        GEMM 1   buf[:, :F] = relu(buf[:, F:] @ W1 + b1)            (bias+ReLU epilogue, strided output)
        GEMM 2   y          = relu(buf @ [Wh ; W_in Wh] + b2)       ([B, F+96] x [F+96, 2F], epilogue)

    and no tensor is touched by a stand-alone elementwise kernel before the two tiny output layers.
    The engine writes its planes straight into columns F..F+84 of `buf` (plane_stride = F + 96,
    plane_offset = F), so the network input needs no copy either.
    """

    def __init__(self, model: ConnectFourNet, dtype: torch.dtype = torch.bfloat16, device=None):
        super().__init__(model, dtype=dtype, device=device)
        if not self.supports(model):
            raise ValueError("FusedNet needs n_residual_blocks == 1 and hidden layers in both heads")
        self.plane_stride = self.F + self.IN_PAD
        self.plane_offset = self.F
        self._refuse()
        self._strided_out_ok = None

    @staticmethod
    def supports(model: ConnectFourNet) -> bool:
        c = model.config
        return c.n_residual_blocks == 1 and c.n_policy_layers >= 2 and c.n_value_layers >= 2

    @torch.no_grad()
    def _refuse(self):
        F = self.F
        w0, b0 = self.w0.double(), self.b0.double()      # [96, 2F]: pre | inp
        wh, bh = self.wh0.double(), self.bh0.double()    # [F, 2F]
        w1, b1 = w0[:, :F], b0[:F]
        w_in, b_in = w0[:, F:], b0[F:]
        w2 = torch.cat([wh, w_in @ wh], dim=0)           # [F + 96, 2F]
        b2 = bh + b_in @ wh
        for name, t in (("f_w1", w1), ("f_b1", b1), ("f_w2", w2), ("f_b2", b2)):
            t = t.to(self.dtype).contiguous()
            if hasattr(self, name):
                getattr(self, name).copy_(t)
            else:
                self.register_buffer(name, t, persistent=False)

    @torch.no_grad()
    def refresh(self, model: ConnectFourNet) -> "FusedNet":
        # NOTE: w0/wh0 are stored in `dtype`; re-fold from the float64 source for full accuracy
        super().refresh(model)
        tensors, _ = self._fold(model)
        F = self.F
        w0, b0, wh, bh = tensors["w0"], tensors["b0"], tensors["wh0"], tensors["bh0"]
        self.f_w1.copy_(w0[:, :F])
        self.f_b1.copy_(b0[:F])
        self.f_w2.copy_(torch.cat([wh, w0[:, F:] @ wh], dim=0))
        self.f_b2.copy_(bh + b0[F:] @ wh)
        return self

    def forward(self, buf: torch.Tensor, out=None):
        """buf: [B, F + 96]; columns F..F+84 hold the input planes (rest of the tail zero).  Columns
        0..F are overwritten.  A [B,2,6,7] tensor is accepted too (copied into a fresh buffer)."""
        F = self.F
        if buf.dim() == 4:
            planes = buf
            buf = planes.new_zeros(planes.shape[0], F + self.IN_PAD)
            buf[:, F : F + 84] = planes.reshape(planes.shape[0], 84)
        x0 = buf[:, F:]
        if buf.is_cuda:
            if self._strided_out_ok is None:
                self._strided_out_ok = self._probe_strided_out(buf)
            if self._strided_out_ok:
                torch._addmm_activation(self.f_b1, x0, self.f_w1, out=buf[:, :F])
            else:
                buf[:, :F] = torch._addmm_activation(self.f_b1, x0, self.f_w1)
            y = torch._addmm_activation(self.f_b2, buf, self.f_w2)
        else:
            buf[:, :F] = torch.relu(torch.addmm(self.f_b1, x0, self.f_w1))
            y = torch.relu(torch.addmm(self.f_b2, buf, self.f_w2))
        hp, hv = y[:, :F], y[:, F:]
        for i in range(self.n_p):
            hp = self._lin_relu(hp, getattr(self, f"wp{i}"), getattr(self, f"bp{i}"))
        for i in range(self.n_v):
            hv = self._lin_relu(hv, getattr(self, f"wv{i}"), getattr(self, f"bv{i}"))
        return self._heads(hp, hv, out)

    def _probe_strided_out(self, buf: torch.Tensor) -> bool:
        """Does the cuBLASLt epilogue path accept a row-strided `out`?  Checked once, numerically."""
        try:
            F = self.F
            t = torch.zeros(8, F + self.IN_PAD, dtype=buf.dtype, device=buf.device)
            t[:, F : F + 84] = (torch.arange(8 * 84, device=buf.device).reshape(8, 84) % 3 == 0).to(buf.dtype)
            want = torch.relu(torch.addmm(self.f_b1, t[:, F:], self.f_w1))
            torch._addmm_activation(self.f_b1, t[:, F:], self.f_w1, out=t[:, :F])
            return bool(torch.allclose(t[:, :F].float(), want.float(), atol=2e-2, rtol=2e-2))
        except Exception:
            return False
