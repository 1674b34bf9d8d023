"""Network graph time per batch size for the inference form the bench runs (FusedNet bf16 with the
fused output kernel), CUDA graphs, CUDA events."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from c4a0_b200.nn import ConnectFourNet, FusedNet, default_config  # noqa: E402

if "--tunable" in sys.argv:  # let PyTorch's TunableOp pick the GEMM algorithm per shape (tuned on first use)
    import time
    import torch.cuda.tunable as tunable

    tunable.enable(True)
    tunable.tuning_enable(True)
    tunable.set_max_tuning_duration(30)
    tunable.set_max_tuning_iterations(20)
    tunable.set_filename("/tmp/tunable.csv")
torch.manual_seed(1337)
model = ConnectFourNet(default_config()).cuda().eval()
net = FusedNet(model, dtype=torch.bfloat16)
if "--gemm-heads" in sys.argv:  # the two output layers as cuBLASLt GEMMs + c4a0_head_epilogue instead of c4a0_heads
    from c4a0_b200.nn import _output_stage

    net._heads = lambda hp, hv, out=None: _output_stage(torch.addmm(net.bpf, hp, net.wpf), torch.addmm(net.bvf, hv, net.wvf), out)
sizes = (64, 128, 256, 512, 1024, 1536, 2048, 3072, 4096, 5120, 6144, 7168, 8192, 10240, 12288, 14336, 16384)
if "--few" in sys.argv:
    sizes = (64, 1024, 4096, 8192, 16384)
flops = model.flops_per_position()
s = torch.cuda.Stream()
for B in sizes:
    buf = torch.zeros(B, net.plane_stride, device="cuda", dtype=torch.bfloat16)
    out = (torch.zeros(B, 7, device="cuda"), torch.zeros(B, device="cuda"), torch.zeros(B, device="cuda"))
    import time as _t
    t0 = _t.time()
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(3):
            net(buf, out=out)
        s.synchronize()
        warm = _t.time() - t0
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            net(buf, out=out)
        g.replay()
        s.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(200):
            g.replay()
        b.record(s)
        s.synchronize()
    us = a.elapsed_time(b) / 200 * 1e3
    print(f"B={B:6d} {us:8.1f} us  {us / B * 1e3:7.1f} ns/row  {B * flops / us / 1e6:7.1f} TFLOP/s (reference-form flops)  warm-up {warm:.2f} s", flush=True)
