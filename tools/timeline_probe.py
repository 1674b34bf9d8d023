"""Timeline of the bench job through the Python loop: every N ticks the live games, rows, and the
counters' increments (simulations, cache hits, rows asked for, speculative rows, terminal sims)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from c4a0_b200.nn import ConnectFourNet, default_config  # noqa: E402
from c4a0_b200.selfplay import DeviceEvaluator, SelfPlaySession  # noqa: E402

n, sims, every = 16384, 600, int(sys.argv[1]) if len(sys.argv) > 1 else 250
torch.manual_seed(1337)
model = ConnectFourNet(default_config()).cuda().eval()
ev = DeviceEvaluator.from_model(model, torch.bfloat16)
sess = SelfPlaySession(n, n, sims, 6.6, 0.01, plane_dtype=torch.bfloat16, plane_stride=ev.plane_stride,
                       plane_offset=ev.plane_offset, n_lanes=1, eval_cache=True, speculate="--no-spec" not in sys.argv)
ln = sess.lanes[0]
ids = np.arange(n)
z = np.zeros(n, np.uint64)
keys = ("sims", "cache_hits", "leaf_requests", "nn_evals", "spec_rows", "terminal_leaf_sims", "moves")
with torch.cuda.stream(ln.stream):
    s = ln.stream.cuda_stream
    ln.engine.set_requests(ids, z, z, s)
    prev = {k: 0 for k in keys}
    t_prev = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ln.stream)
    for tick in range(1, 40001):
        rows = ln.engine.poll(s).n_rows if tick % 8 == 1 else rows
        b = min(ln.io_rows, max(128, ((rows + 511) // 512) * 512 + 1024))
        ln.evaluate(ev, b)
        ln.engine.step(s)
        if tick % every == 0:
            e1.record(ln.stream)
            ln.stream.synchronize()
            ms = e0.elapsed_time(e1)
            p = ln.engine.poll(s)
            st = ln.engine.stats(s)
            d = {k: st[k] - prev[k] for k in keys}
            live = max(1, p.n_running)
            print(f"tick {tick:6d} live {p.n_running:6d} rows {p.n_rows:6d} | per tick: sims/game {d['sims'] / every / live:5.2f} "
                  f"hits {d['cache_hits'] / every:8.0f} asked {d['leaf_requests'] / every:8.0f} rows {d['nn_evals'] / every:8.0f} "
                  f"spec {d['spec_rows'] / every:8.0f} term {d['terminal_leaf_sims'] / every:8.0f} moves {d['moves'] / every:7.1f} "
                  f"| {ms / every * 1e3:7.1f} us/tick (python loop)", flush=True)
            prev = {k: st[k] for k in keys}
            if p.n_finished == n:
                break
            e0.record(ln.stream)
