"""Target for ncu: runs the tree kernels alone (synthetic hash evaluator) at the bench workload so
that steady-state launches of k_step / k_tail can be captured.

    ncu ... python tools/profile_target.py --games 16384 --sims 600 --ticks 2600
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from c4a0_b200 import _lib as L  # noqa: E402
from c4a0_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=16384)
ap.add_argument("--sims", type=int, default=600)
ap.add_argument("--ticks", type=int, default=2600)
ap.add_argument("--bf16", action="store_true")
ap.add_argument("--no-dedup", action="store_true")
ap.add_argument("--cache", action="store_true", help="evaluation cache on (as the bench runs)")
ap.add_argument("--flat", action="store_true", help="near-uniform hash evaluator (a freshly initialised network)")
ap.add_argument("--arena-blocks", type=int, default=0)
a = ap.parse_args()
n = a.games
e = Engine(n, n, a.sims, 6.6, 0.01, L.PLANES_BF16 if a.bf16 else L.PLANES_F32, 0, 0, 96,
           (L.FLAG_NO_DEDUP if a.no_dedup else 0) | (L.FLAG_EVAL_CACHE if a.cache else 0), a.arena_blocks)
planes = torch.zeros(n, 96, device="cuda", dtype=torch.bfloat16 if a.bf16 else torch.float32)
logits = torch.zeros(n, 7, device="cuda")
qp = torch.zeros(n, device="cuda")
qn = torch.zeros(n, device="cuda")
e.bind_io(planes.data_ptr(), logits.data_ptr(), qp.data_ptr(), qn.data_ptr())
e.set_requests(list(range(n)), [0] * n, [0] * n)
for t in range(a.ticks):
    e.eval_builtin(L.EVAL_HASH_FLAT if a.flat else L.EVAL_HASH)
    e.step()
p = e.poll()
st = e.stats()
print("ticks", a.ticks, "finished", p.n_finished, "rows", p.n_rows, "stats", st)
