"""The c4a0 network on the library's own sm_100a kernel (include/c4a0_net.h, csrc/net.cu).

`NativeEvaluator(model)` folds a `ConnectFourNet` (same weights, same function as the reference's
src/c4a0/nn.py:59-117 in eval mode) into the dense-layer program the kernel runs:

    buffer 0  [ H1 (Fp) | planes (128) ]     L1: H1 = relu(planes @ W1 + b1)             (stem + residual block pre-activation)
    buffer 1  [ hp (Fp) | hv (Fp) ]          L2: [hp | hv] = relu([H1 | planes] @ W2 + b2) (block output folded in; first
                                                  hidden layer of both heads, nn.py:75-100)
    buffers 2.. ping-pong [Fp]               further hidden layers of the policy / value head
    output layers (16 wide)                  log_softmax(7) -> logits, tanh(2) -> q_penalty, q_no_penalty

(the algebra is FusedNet's, c4a0_b200/nn.py).  Fp is F = 42 * conv_filter_size rounded up to a multiple
of 1344 (a common multiple of the kernel's column tiles); padding columns carry zero weights and zero biases.  Weights live in
torch tensors owned by this object; `refresh(model)` overwrites them in place for a new generation.
Torch is used for the folding arithmetic (float64, once per generation) and for device memory only.
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib as L
from .nn import ConnectFourNet, FoldedNet

XP = 128  # columns reserved for the 84 input planes (a multiple of the 64-wide K step)


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@torch.no_grad()
def fold_program(model: ConnectFourNet):
    """The network as the kernel's layer program: ([(name, W [n_pad, k_pad] f64, b [n_pad] f64, meta)], n_buffers),
    meta = dict(kind, inp=(buffer, col0), outp=(buffer, col0), dep=layer index or -1).  Pure torch (any device)."""
    t, (n_blocks, joint_first, n_p, n_v) = FoldedNet._fold(model)
    assert n_blocks == 1 and joint_first
    F = model.fc_size
    Fp = _round_up(F, L.NET_PAD_N)
    dev = t["w0"].device
    z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)  # noqa: E731
    w0, b0, wh, bh = t["w0"], t["b0"], t["wh0"], t["bh0"]  # [96,2F] (pre|inp), [F,2F] (policy|value)
    w1, b1, w_in, b_in = w0[:84, :F], b0[:F], w0[:84, F:], b0[F:]
    out = []
    # L1: [Fp][128]
    W = z(Fp, XP)
    W[:F, :84] = w1.t()
    b = z(Fp)
    b[:F] = b1
    out.append(("l1", W, b, dict(kind=L.NET_HIDDEN, inp=(0, Fp), outp=(0, 0), dep=-1)))
    # L2: [2Fp][Fp+128]; h = relu(pre) + inp, inp = planes @ w_in + b_in
    #     =>  h @ wh = relu(pre) @ wh + planes @ (w_in wh) + b_in wh
    W = z(2 * Fp, Fp + XP)
    x = w_in @ wh  # [84, 2F]
    W[:F, :F] = wh[:, :F].t()
    W[Fp : Fp + F, :F] = wh[:, F:].t()
    W[:F, Fp : Fp + 84] = x[:, :F].t()
    W[Fp : Fp + F, Fp : Fp + 84] = x[:, F:].t()
    b = z(2 * Fp)
    b2 = bh + b_in @ wh
    b[:F] = b2[:F]
    b[Fp : Fp + F] = b2[F:]
    out.append(("l2", W, b, dict(kind=L.NET_HIDDEN, inp=(0, 0), outp=(1, 0), dep=0)))

    def hidden(name, w, bias):
        W = z(Fp, Fp)
        W[:F, :F] = w.t()
        b = z(Fp)
        b[:F] = bias
        return name, W, b

    def head(name, w, bias, n_out):
        W = z(L.NET_HEAD_N, Fp)
        W[:n_out, :F] = w[:, :n_out].t()
        b = z(L.NET_HEAD_N)
        b[:n_out] = bias[:n_out]
        return name, W, b

    # policy chain reads buffer 1 columns [0, Fp), value chain columns [Fp, 2 Fp); ping-pong buffers after that
    nxt = [2]

    def chain(prefix, n_hidden, first_in, kind, wf, bf, n_out):
        layers, inp, dep = [], first_in, "l2"
        for i in range(n_hidden):
            buf = nxt[0]
            nxt[0] += 1
            nm, W, b = hidden(f"{prefix}{i}", t[f"w{prefix}{i}"], t[f"b{prefix}{i}"])
            layers.append([nm, W, b, dict(kind=L.NET_HIDDEN, inp=inp, outp=(buf, 0), dep=dep)])
            inp, dep = (buf, 0), nm
        nm, W, b = head(f"{prefix}f", wf, bf, n_out)
        layers.append([nm, W, b, dict(kind=kind, inp=inp, outp=(0, 0), dep=dep)])
        return layers

    pol = chain("p", n_p, (1, 0), L.NET_POLICY, t["wpf"], t["bpf"], 7)
    val = chain("v", n_v, (1, Fp), L.NET_VALUE, t["wvf"], t["bvf"], 2)
    # program order: one policy layer first (so L2 is complete for most row tiles when the value chain's
    # tiles come up), then the value chain, then the rest of the policy chain
    seq = pol[:1] + val + pol[1:] if len(pol) > 1 else val + pol
    names = ["l1", "l2"] + [s[0] for s in seq]
    for s in seq:
        s[3]["dep"] = names.index(s[3]["dep"])
        out.append(tuple(s))
    return out, nxt[0]


def emulate_program(layers, n_buffers, planes: torch.Tensor, dtype=torch.float64):
    """Reference semantics of the layer program in plain torch (tests): planes [B,84] -> (logits, qp, qn).
    With dtype=bfloat16 weights and stored activations are rounded like the kernel's (f32 accumulation)."""
    B = planes.shape[0]
    cols = {}
    for name, W, b, m in layers:
        for buf, c0, width in ((m["inp"][0], m["inp"][1], W.shape[1]), (m["outp"][0], m["outp"][1], W.shape[0] if m["kind"] == L.NET_HIDDEN else 0)):
            cols[buf] = max(cols.get(buf, 0), c0 + width)
    acc = torch.float32 if dtype == torch.bfloat16 else dtype
    bufs = {k: torch.zeros(B, v, dtype=acc, device=planes.device) for k, v in cols.items()}
    p_buf, p_col = layers[0][3]["inp"]
    bufs[p_buf][:, p_col : p_col + 84] = planes.reshape(B, 84).to(acc)
    logits = qp = qn = None
    for name, W, b, m in layers:
        Wq = W.to(dtype).to(acc)
        x = bufs[m["inp"][0]][:, m["inp"][1] : m["inp"][1] + W.shape[1]]
        y = x @ Wq.t() + b.to(acc)
        if m["kind"] == L.NET_HIDDEN:
            y = torch.relu(y)
            if dtype == torch.bfloat16:
                y = y.to(torch.bfloat16).to(acc)
            bufs[m["outp"][0]][:, m["outp"][1] : m["outp"][1] + W.shape[0]] = y
        elif m["kind"] == L.NET_POLICY:
            logits = torch.log_softmax(y[:, :7], dim=1)
        else:
            q = torch.tanh(y[:, :2])
            qp, qn = q[:, 0], q[:, 1]
    return logits, qp, qn


class _DevArray:
    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class NativeEvaluator:
    """Folded weights of one model on one device, in the layout of the kernel's layer program."""

    dtype = torch.bfloat16

    def __init__(self, model: ConnectFourNet, device=None):
        if not self.supports(model):
            raise ValueError("the native network kernel covers n_residual_blocks == 1 with hidden layers in both heads")
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("NativeEvaluator needs a CUDA device: there is no CPU fallback")
        self.F = model.fc_size
        self.Fp = _round_up(self.F, L.NET_PAD_N)
        self.plane_stride = self.Fp + XP
        self.plane_offset = self.Fp
        self._layers: List[dict] = []
        self._build(model)

    @staticmethod
    def supports(model) -> bool:
        if not isinstance(model, ConnectFourNet):
            return False
        c = model.config
        return c.n_residual_blocks == 1 and c.n_policy_layers >= 2 and c.n_value_layers >= 2

    def _fold(self, model: ConnectFourNet):
        layers, n_buffers = fold_program(model)
        self.n_buffers = n_buffers
        return layers

    @torch.no_grad()
    def _build(self, model: ConnectFourNet):
        folded = self._fold(model)
        self._layers = []
        for name, W, b, meta in folded:
            self._layers.append(dict(
                name=name, meta=meta,
                w=W.to(device=self.device, dtype=torch.bfloat16).contiguous(),
                b=b.to(device=self.device, dtype=torch.float32).contiguous(),
            ))
        self.buffer_cols = [self.Fp + XP, 2 * self.Fp] + [self.Fp] * (self.n_buffers - 2)

    @torch.no_grad()
    def refresh(self, model: ConnectFourNet) -> "NativeEvaluator":
        """A new generation's weights, in place: nets created from this evaluator keep working."""
        if not self.supports(model) or model.fc_size != self.F:
            raise ValueError("refresh() needs a model of the same architecture")
        folded = self._fold(model)
        if [f[0] for f in folded] != [l["name"] for l in self._layers]:
            raise ValueError("refresh() needs a model of the same architecture")
        for (name, W, b, meta), cur in zip(folded, self._layers):
            cur["w"].copy_(W)
            cur["b"].copy_(b)
        return self

    def flops_per_row(self) -> int:
        return sum(2 * l["w"].shape[0] * l["w"].shape[1] for l in self._layers)

    # ---- instances -------------------------------------------------------------------------------
    def instantiate(self, max_rows: int) -> "NativeNet":
        return NativeNet(self, max_rows)

    def __call__(self, planes: torch.Tensor):
        """planes [B,2,6,7] / [B,84] -> (logits [B,7], q_penalty [B], q_no_penalty [B]) float32 CUDA tensors
        (copies), through a private net sized for the largest batch seen.  The self-play engine does not go
        through here: every lane owns a net whose input buffer the engine writes directly."""
        B = int(planes.shape[0])
        net = getattr(self, "_own_net", None)
        if net is None or net.max_rows < B:
            if net is not None:
                net.close()
            net = self._own_net = self.instantiate(max(B, 256))
        return tuple(t.clone() for t in net(planes.to(self.device)))


class NativeNet:
    """One c4a0_net: activation buffers for up to max_rows rows over an evaluator's weights."""

    def __init__(self, ev: NativeEvaluator, max_rows: int):
        self.ev = ev
        self.max_rows = int(max_rows)
        spec = L.NetSpec()
        spec.device = ev.device.index if ev.device.index is not None else torch.cuda.current_device()
        spec.max_rows = self.max_rows
        spec.n_buffers = ev.n_buffers
        for i, c in enumerate(ev.buffer_cols):
            spec.buffer_cols[i] = c
        spec.planes_buffer, spec.planes_col0 = 0, ev.plane_offset
        spec.n_layers = len(ev._layers)
        for i, l in enumerate(ev._layers):
            m = l["meta"]
            sl = spec.layers[i]
            sl.weight_dev, sl.bias_dev = l["w"].data_ptr(), l["b"].data_ptr()
            sl.n_pad, sl.k_pad = l["w"].shape
            sl.in_buffer, sl.in_col0 = m["inp"]
            sl.out_buffer, sl.out_col0 = m["outp"]
            sl.dep, sl.kind = m["dep"], m["kind"]
        self._lib = L.lib()
        self._h = C.c_void_p()
        L.check(self._lib.c4a0_net_create(C.byref(spec), C.byref(self._h)))
        self._outputs = None

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.c4a0_net_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self) -> int:
        return int(self._lib.c4a0_net_device_bytes(self._h))

    def buffer(self, index: int) -> torch.Tensor:
        """Activation buffer `index` as a bf16 tensor [rows, cols] over the library's memory (no copy)."""
        base, cols, rows = C.c_void_p(), C.c_uint32(), C.c_uint32()
        L.check(self._lib.c4a0_net_buffer(self._h, index, C.byref(base), C.byref(cols), C.byref(rows)))
        raw = torch.as_tensor(_DevArray(base.value, (rows.value * cols.value,), "<i2"), device=self.ev.device)
        return raw.view(torch.bfloat16).view(rows.value, cols.value)

    def planes_ptr(self) -> int:
        """Device address of row 0's first plane element (what c4a0_engine_bind_io gets)."""
        base = C.c_void_p()
        L.check(self._lib.c4a0_net_buffer(self._h, 0, C.byref(base), None, None))
        return base.value + 2 * self.ev.plane_offset

    def bind_outputs(self, logits: torch.Tensor, qp: torch.Tensor, qn: torch.Tensor) -> None:
        for t, n in ((logits, 7 * self.max_rows), (qp, self.max_rows), (qn, self.max_rows)):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() >= n):
                raise ValueError("outputs must be contiguous float32 CUDA tensors with max_rows rows")
        self._outputs = (logits, qp, qn)
        L.check(self._lib.c4a0_net_bind_outputs(self._h, logits.data_ptr(), qp.data_ptr(), qn.data_ptr()))

    def bind_row_count(self, a_ptr: Optional[int], b_ptr: Optional[int]) -> None:
        L.check(self._lib.c4a0_net_bind_row_count(self._h, a_ptr, b_ptr))

    def forward(self, rows: int, stream: Optional[int] = None) -> None:
        s = torch.cuda.current_stream(self.ev.device).cuda_stream if stream is None else stream
        L.check(self._lib.c4a0_net_forward(self._h, rows, s))

    def forward_timed(self, rows: int, stream: Optional[int] = None) -> float:
        s = torch.cuda.current_stream(self.ev.device).cuda_stream if stream is None else stream
        ms = C.c_float()
        L.check(self._lib.c4a0_net_forward_timed(self._h, rows, s, C.byref(ms)))
        return ms.value

    def debug_trace(self, rows: int, cta: int = 0):
        """Per-role event log of one CTA: three lists of (tag, SM cycle) (see c4a0_net_debug_trace)."""
        import numpy as np

        out = np.zeros((3, 4096), np.uint64)
        s = torch.cuda.current_stream(self.ev.device).cuda_stream
        L.check(self._lib.c4a0_net_debug_trace(self._h, rows, cta, s, L.ptr(out), out.size))
        return [[(int(v) & 0xFF, int(v) >> 8) for v in role if v] for role in out]

    def __call__(self, planes: torch.Tensor):
        """planes [B,2,6,7] or [B,84] (any float dtype) -> (logits [B,7], q_penalty [B], q_no_penalty [B]) f32.
        Convenience for tests: copies the planes in, runs the kernel, returns views of the bound outputs."""
        B = planes.shape[0]
        if self._outputs is None:
            self.bind_outputs(torch.zeros(self.max_rows, 7, device=self.ev.device), torch.zeros(self.max_rows, device=self.ev.device),
                              torch.zeros(self.max_rows, device=self.ev.device))
        buf = self.buffer(0)
        o = self.ev.plane_offset
        buf[:B, o : o + 84] = planes.reshape(B, 84).to(torch.bfloat16)
        self.forward(B)
        lg, qp, qn = self._outputs
        return lg[:B], qp[:B], qn[:B]
