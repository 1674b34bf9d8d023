"""Simulations, live games and rows of given ticks of the default bench job (python loop), so that an ncu capture
of 'the k_step launch of tick T' can be set against that launch's own algorithmic bytes."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from c4a0_b200.nn import ConnectFourNet, default_config  # noqa: E402
from c4a0_b200.selfplay import DeviceEvaluator, SelfPlaySession  # noqa: E402

ticks = [int(x) for x in sys.argv[1].split(",")]
n, sims = 16384, 600
torch.manual_seed(1337)
model = ConnectFourNet(default_config()).cuda().eval()
ev = DeviceEvaluator.from_model(model, torch.bfloat16)
sess = SelfPlaySession(n, n, sims, 6.6, 0.01, plane_dtype=torch.bfloat16, plane_stride=ev.plane_stride,
                       plane_offset=ev.plane_offset, n_lanes=1, eval_cache=True, speculate=True)
ln = sess.lanes[0]
ln.attach_native(ev)
out = {}
with torch.cuda.stream(ln.stream):
    s = ln.stream.cuda_stream
    ln.engine.set_requests(np.arange(n), np.zeros(n, np.uint64), np.zeros(n, np.uint64), s)
    for t in range(1, max(ticks) + 1):  # tick t = the t-th k_step launch of the job
        ln.evaluate(ev, ln.io_rows)
        if t in ticks:
            a = ln.engine.stats(s)
            rows_in = ln.engine.poll(s).n_rows
        ln.engine.step(s)
        if t in ticks:
            b, p = ln.engine.stats(s), ln.engine.poll(s)
            out[t] = dict(sims=b["sims"] - a["sims"], depth_sum=b["select_depth_sum"] - a["select_depth_sum"],
                          expansions=b["expansions"] - a["expansions"], live=p.n_running, rows_evaluated_before=rows_in, rows_packed=p.n_rows)
print(json.dumps(out))
