"""How much of a k_step launch is its tail?  Plays the bench job (16,384 games x 600 sims, native
bf16 network, cache + speculation) through the Python loop and, at a few ticks along the job, runs one
tick with the per-game cycle counters (c4a0_engine_debug_phases): the distribution over live games of
the time their warp needed, against the duration of the whole launch, and the network time at that
tick's row count."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from c4a0_b200.nn import ConnectFourNet, default_config  # noqa: E402
from c4a0_b200.selfplay import DeviceEvaluator, SelfPlaySession  # noqa: E402

n, sims = 16384, 600
at = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [100, 400, 1000, 2000, 3000, 4000, 5000, 6000, 7000, 7500]
torch.manual_seed(1337)
model = ConnectFourNet(default_config()).cuda().eval()
ev = DeviceEvaluator.from_model(model, torch.bfloat16)
sess = SelfPlaySession(n, n, sims, 6.6, 0.01, plane_dtype=torch.bfloat16, plane_stride=ev.plane_stride,
                       plane_offset=ev.plane_offset, n_lanes=1, eval_cache=True, speculate=True)
ln = sess.lanes[0]
if hasattr(ev, "instantiate"):
    ln.attach_native(ev)
ids = np.arange(n)
z = np.zeros(n, np.uint64)
MHZ = 1965.0
with torch.cuda.stream(ln.stream):
    s = ln.stream.cuda_stream
    ln.engine.set_requests(ids, z, z, s)
    for tick in range(1, 40001):
        ln.evaluate(ev, ln.io_rows)
        if tick in at:
            p0 = ln.engine.poll(s)
            d = ln.engine.debug_phases(s).astype(np.int64)
            act = d[:, 7] > 0
            a = d[act]
            p = ln.engine.poll(s)
            tot = a[:, 7] / MHZ
            pc = np.percentile(tot, [50, 75, 90, 95, 99, 99.9, 100])
            print(f"tick {tick}: live {p.n_running} rows in {p0.n_rows} -> out {p.n_rows}; games stepped {act.sum()}")
            print("   per-game warp time us  p50 %.1f p75 %.1f p90 %.1f p95 %.1f p99 %.1f p99.9 %.1f max %.1f" % tuple(pc))
            sm = a[:, 2]
            print("   sims/game: mean %.2f, histogram %s" % (sm.mean(), np.bincount(np.minimum(sm, 16)).tolist()))
            for k in (1, 2, 3, 4, 6, 8, 12):
                sel = sm == k
                if sel.sum():
                    print(f"      games with {k} sims: {sel.sum()}, mean warp time {tot[sel].mean():.1f} us")
            ln.evaluate(ev, ln.io_rows)
            x, y = ln.engine.step_timed(s)
            p2 = ln.engine.poll(s)
            t_nn = ln.net.forward_timed(p2.n_rows, s) * 1e3 if ln.net is not None else float("nan")
            print("   next tick: k_step %.1f us (k_tail %.1f us), then network on %d rows %.1f us" % (x * 1e3, y * 1e3, p2.n_rows, t_nn), flush=True)
            ln.engine.step(s)
        else:
            ln.engine.step(s)
        if tick % 250 == 0 and ln.engine.poll(s).n_finished == n:
            print("finished at tick", tick)
            break
