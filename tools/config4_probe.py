"""Config 4 of BASELINE.json at reduced length: 131,072 concurrent games x 1,600 sims/move, wider
ResNet (conv_filter_size 64), bf16.  Runs a fixed number of ticks through the Python loop and
reports tick rate, memory and tree-kernel time — a capacity / throughput probe, not the bench."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from c4a0_b200.nn import ConnectFourNet, ModelConfig  # noqa: E402
from c4a0_b200.selfplay import DeviceEvaluator, SelfPlaySession  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 2400
width = int(sys.argv[4]) if len(sys.argv) > 4 else 64
torch.manual_seed(1337)
model = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=width, n_policy_layers=4, n_value_layers=2)).cuda().eval()
ev = DeviceEvaluator.from_model(model, torch.bfloat16)
t0 = time.time()
sess = SelfPlaySession(n, n, sims, 6.6, 0.01, plane_dtype=torch.bfloat16, plane_stride=ev.plane_stride,
                       plane_offset=ev.plane_offset, n_lanes=1, eval_cache="--cache" in sys.argv)
ln = sess.lanes[0]
print("engine bytes %.1f GB, create %.1f s, arena_blocks/half %d" % (ln.engine.device_bytes / 1e9, time.time() - t0, ln.engine.cfg.arena_blocks), flush=True)
ids = np.arange(n)
z = np.zeros(n, np.uint64)
ks = km = 0.0
kn = 0
with torch.cuda.stream(ln.stream):
    s = ln.stream.cuda_stream
    ln.engine.set_requests(ids, z, z, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ln.stream)
    for t in range(ticks):
        rows = ln.engine.poll(s).n_rows if t % 50 == 0 else rows
        b = min(n, max(128, ((rows + 4095) // 4096) * 4096 + (4096 if t % 50 else 8192)))  # generous cover of the live rows
        ln.evaluate(ev, min(n, b))
        if t % 100 == 99:
            x, y = ln.engine.step_timed(s)
            ks, km, kn = ks + x, km + y, kn + 1
        else:
            ln.engine.step(s)
    e1.record(ln.stream)
    ln.stream.synchronize()
    st = ln.engine.stats(s)
    p = ln.engine.poll(s)
ms = e0.elapsed_time(e1)
d = st["select_depth_sum"] / max(1, st["sims"])
e = st["expansions"] / max(1, st["sims"])
bps = 92 * d + 24 * (d + 1) + 16 + 84 * 2 + 36 + 168 * e
spl = st["sims"] / ticks
print({"ticks": ticks, "ms_per_tick": ms / ticks, "sims": st["sims"], "sims_per_s": st["sims"] / ms * 1e3, "rows_now": p.n_rows,
       "finished": p.n_finished, "k_step_ms": ks / kn, "k_tail_ms": km / kn, "depth": d, "bytes_per_sim": bps,
       "k_step_GBps": bps * spl / (ks / kn * 1e-3) / 1e9, "compactions": st["compactions"]})
print("max mem allocated by torch %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))
