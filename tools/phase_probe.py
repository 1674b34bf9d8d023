"""Where does a tick of k_step spend its cycles?  Runs the bench workload with the real (folded,
bf16) network through the Python loop and samples per-game phase cycle counters at several points."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from c4a0_b200.nn import ConnectFourNet, default_config  # noqa: E402
from c4a0_b200.selfplay import DeviceEvaluator, SelfPlaySession  # noqa: E402

n, sims = int(sys.argv[1]) if len(sys.argv) > 1 else 16384, 600
torch.manual_seed(1337)
model = ConnectFourNet(default_config()).cuda().eval()
ev = DeviceEvaluator.from_model(model, torch.bfloat16)
sess = SelfPlaySession(n, n, sims, 6.6, 0.01, plane_dtype=torch.bfloat16, plane_stride=ev.plane_stride,
                       plane_offset=ev.plane_offset, n_lanes=1, eval_cache="--no-cache" not in sys.argv)
ln = sess.lanes[0]
ids = np.arange(n)
z = np.zeros(n, np.uint64)
names = ["load", "fetch_answer", "sims", "run", "inline_sims", "depth", "store", "total"]
with torch.cuda.stream(ln.stream):
    s = ln.stream.cuda_stream
    ln.engine.set_requests(ids, z, z, s)
    for tick in range(1, 20001):
        ln.evaluate(ev, ln.io_rows)
        if tick in (300, 700, 1500, 3000, 5000, 7000, 9000, 10500):
            d = ln.engine.debug_phases(s).astype(np.int64)
            act = d[:, 7] > 0
            a = d[act]
            p = ln.engine.poll(s)
            print(f"tick {tick}: active {act.sum()} rows {p.n_rows} finished {p.n_finished}")
            print("   mean:", {k: round(float(a[:, i].mean()), 1) for i, k in enumerate(names)})
            print("   max :", {k: int(a[:, i].max()) for i, k in enumerate(names)})
            print("   slowest game: %.1f us at 1.9 GHz" % (a[:, 7].max() / 1900))
            x, y = ln.engine.step_timed(s)
            ln.evaluate(ev, ln.io_rows)
            print("   step_timed: k_step %.1f us k_tail %.1f us" % (x * 1e3, y * 1e3))
        else:
            ln.engine.step(s)
        if tick % 500 == 0 and ln.engine.poll(s).n_finished == n:
            break
