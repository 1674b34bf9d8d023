"""Numerics (per buffer) and timing of the native network kernel against plain torch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch
from c4a0_b200 import _lib as L
from c4a0_b200.native_net import NativeEvaluator, emulate_program, fold_program
from c4a0_b200.nn import ConnectFourNet, ModelConfig, FusedNet

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=32)
ap.add_argument("--rows", type=int, nargs="*", default=[1, 128, 300, 1024, 5120, 16384])
ap.add_argument("--check-rows", type=int, default=300)
ap.add_argument("--iters", type=int, default=30)
ap.add_argument("--no-time", action="store_true")
ap.add_argument("--configs", nargs="*", default=["pair", "pair:1", "pair:2", "pair:3", "one"],
                help="pair = CTA-pair kernel with automatic column tile, pair:N = forced variant N (1: 224, 2: 96, 3: 32), one = one-CTA kernel")
a = ap.parse_args()

torch.manual_seed(1337)
dev = torch.device("cuda", 0)
model = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=a.width, n_policy_layers=4, n_value_layers=2)).to(dev).eval()
ev = NativeEvaluator(model)
BODY = r"""
cap = max(a.rows + [a.check_rows])
os.environ.pop("C4A0_NET_ONE_CTA", None); os.environ.pop("C4A0_NET_VARIANT", None)
if CONFIG == "one":
    os.environ["C4A0_NET_ONE_CTA"] = "1"
elif ":" in CONFIG:
    os.environ["C4A0_NET_VARIANT"] = CONFIG.split(":")[1]
net = ev.instantiate(cap)
print(f"==== config {CONFIG}", flush=True)
print(f"F={ev.F} Fp={ev.Fp} layers={[l['name'] for l in ev._layers]} buffers={ev.buffer_cols} net bytes={net.device_bytes/1e6:.1f} MB", flush=True)
B = a.check_rows
g = torch.Generator(device="cpu").manual_seed(5)
planes = (torch.rand(B, 84, generator=g) < 0.25).float().to(dev)
lg, qp, qn = net(planes)
torch.cuda.synchronize()
print("forward ok", flush=True)
layers, nb = fold_program(model.double())
model.float()
# bf16-rounded emulation on the GPU
e_lg, e_qp, e_qn = emulate_program([(n, W.to(dev), b.to(dev), m) for n, W, b, m in layers], nb, planes, dtype=torch.bfloat16)
w_lg, w_qp, w_qn = model(planes.view(B, 2, 6, 7))
for name, x, y in (("logits", lg, e_lg), ("qp", qp, e_qp), ("qn", qn, e_qn)):
    print(f"{name}: max|native - bf16 emulation| = {(x - y).abs().max().item():.3e}   vs f32 module = {(x - {'logits': w_lg, 'qp': w_qp, 'qn': w_qn}[name]).abs().max().item():.3e}", flush=True)
# intermediate buffers against a step-by-step emulation
Fp = ev.Fp
acc = {}
bufs = {0: torch.zeros(B, Fp + 128, device=dev), 1: torch.zeros(B, 2 * Fp, device=dev)}
for i in range(2, nb):
    bufs[i] = torch.zeros(B, Fp, device=dev)
bufs[0][:, Fp:Fp + 84] = planes
for n, W, b, m in layers:
    W = W.to(dev).to(torch.bfloat16).float(); b = b.to(dev).float()
    x = bufs[m["inp"][0]][:, m["inp"][1]: m["inp"][1] + W.shape[1]]
    y = x @ W.t() + b
    if m["kind"] == L.NET_HIDDEN:
        y = torch.relu(y).to(torch.bfloat16).float()
        bufs[m["outp"][0]][:, m["outp"][1]: m["outp"][1] + W.shape[0]] = y
        got = net.buffer(m["outp"][0])[:B, m["outp"][1]: m["outp"][1] + W.shape[0]].float()
        d = (got - y).abs()
        print(f"layer {n}: out max|diff| = {d.max().item():.3e}, mean = {d.mean().item():.3e}, ref mean|y| = {y.abs().mean().item():.3e}, bad cols = {(d.max(0).values > 0.05).sum().item()} rows = {(d.max(1).values > 0.05).sum().item()}", flush=True)
# batch invariance: row 7 alone vs in the batch
one = net(planes[7:8])
one = [t.clone() for t in one]
full = net(planes)
print("batch invariant:", all(torch.equal(o[0], f[7]) for o, f in zip(one, full)), flush=True)
if not a.no_time:
    fused = FusedNet(model, dtype=torch.bfloat16)
    for rows in a.rows:
        buf = torch.zeros(rows, fused.F + 96, dtype=torch.bfloat16, device=dev)
        buf[:, fused.F:fused.F + 84] = (torch.rand(rows, 84, device=dev) < 0.25).to(torch.bfloat16)
        out = (torch.zeros(rows, 7, device=dev), torch.zeros(rows, device=dev), torch.zeros(rows, device=dev))
        net.buffer(0)[:rows, Fp:Fp + 84] = buf[:, fused.F:fused.F + 84]
        for _ in range(3):
            net.forward(rows); fused(buf, out=out)
        torch.cuda.synchronize()
        ts = sorted(net.forward_timed(rows) for _ in range(a.iters))
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fused(buf, out=out); s.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=s):
                fused(buf, out=out)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tt = []
            for _ in range(a.iters):
                e0.record(s); gr.replay(); e1.record(s); e1.synchronize(); tt.append(e0.elapsed_time(e1))
        tt.sort()
        fl = ev.flops_per_row() * rows
        print(f"rows {rows:6d}: native {1e3*ts[len(ts)//2]:8.1f} us (min {1e3*ts[0]:.1f})  {fl/ts[len(ts)//2]/1e9:7.1f} TFLOP/s | cuBLASLt graph {1e3*tt[len(tt)//2]:8.1f} us (min {1e3*tt[0]:.1f})", flush=True)

"""
for CONFIG in a.configs:
    try:
        exec(BODY)
    except Exception as exc:
        print(f"config {CONFIG} FAILED: {exc!r}", flush=True)
        break
