"""Times the network forms at several batch sizes (CUDA events, after warm-up, CUDA graphs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from c4a0_b200.nn import ConnectFourNet, FoldedNet, FusedNet, default_config  # noqa: E402

torch.manual_seed(1337)
model = ConnectFourNet(default_config()).cuda().eval()
flops = model.flops_per_position()


def timeit(fn, x, iters=50):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(3):
            fn(x)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn(x)
        g.replay()
        s.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(iters):
            g.replay()
        b.record(s)
        s.synchronize()
    return a.elapsed_time(b) / iters


for B in (128, 1024, 4096, 8192, 16384, 65536):
    row = [f"B={B:6d}"]
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        m = model.to(dt)
        x = torch.zeros(B, 2, 6, 7, device="cuda", dtype=dt)
        ms = timeit(m, x)
        row.append(f"module {name} {ms*1e3:8.1f} us ({B*flops/ms/1e9:7.1f} TF/s)")
        f = FoldedNet(model.float(), dtype=dt)
        xp = torch.zeros(B, 96, device="cuda", dtype=dt)
        ms = timeit(f, xp)
        row.append(f"folded {name} {ms*1e3:8.1f} us ({B*flops/ms/1e9:7.1f} TF/s ref-flops)")
        fu = FusedNet(model.float(), dtype=dt)
        xb = torch.zeros(B, fu.plane_stride, device="cuda", dtype=dt)
        ms = timeit(fu, xb)
        row.append(f"fused {name} {ms*1e3:8.1f} us (strided out {fu._strided_out_ok})")
    print(" | ".join(row), flush=True)
