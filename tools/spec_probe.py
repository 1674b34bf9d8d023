"""How far does speculation get?  A few games in many slots (every batch has spare rows), flat hash
evaluator; prints ticks, hits, speculative rows and simulations per game per tick."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from c4a0_b200 import _lib as L  # noqa: E402
from c4a0_b200.engine import Engine  # noqa: E402

n_games = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_slots = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 600
for name, flags, spec_rows, inl in (("cache", L.FLAG_EVAL_CACHE, 0, 0), ("spec2048", L.FLAG_EVAL_CACHE | L.FLAG_SPECULATE, 2048, 0),
                                    ("spec8192", L.FLAG_EVAL_CACHE | L.FLAG_SPECULATE, 8192, 0),
                                    ("spec8192 inline 16", L.FLAG_EVAL_CACHE | L.FLAG_SPECULATE, 8192, 16)):
    e = Engine(n_slots, n_games, sims, 6.6, 0.01, L.PLANES_BF16, inl, 0, 96, flags, 0, 0, spec_rows)
    R = e.io_rows
    planes = torch.zeros(R, 96, device="cuda", dtype=torch.bfloat16)
    logits = torch.zeros(R, 7, device="cuda")
    qp = torch.zeros(R, device="cuda")
    qn = torch.zeros(R, device="cuda")
    e.bind_io(planes.data_ptr(), logits.data_ptr(), qp.data_ptr(), qn.data_ptr())
    e.set_requests(list(range(n_games)), [0] * n_games, [0] * n_games)
    t = 0
    max_rows = 0
    while True:
        e.eval_builtin(L.EVAL_HASH_FLAT)
        e.step()
        t += 1
        if t % 16 == 0:
            p = e.poll()
            max_rows = max(max_rows, p.n_rows)
            if p.n_finished == n_games:
                break
    st = e.stats()
    print(f"{name:20s} ticks {t:6d} sims {st['sims']} expansions {st['expansions']} hits {st['cache_hits']} "
          f"asked {st['leaf_requests']} rows {st['nn_evals']} spec {st['spec_rows']} max_rows {max_rows} "
          f"sims/game/tick {st['sims'] / n_games / t:.2f}", flush=True)
    e.close()
