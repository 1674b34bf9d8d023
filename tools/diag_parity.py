"""Which engine configuration reproduces the oracle at n = 600 with the flat hash evaluator?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from c4a0_b200 import _lib as L
from c4a0_b200.engine import Engine

def records(soa, i):
    n = int(soa.n_samples[i])
    return [(int(soa.mask[i, k]), int(soa.value[i, k]), tuple(soa.policy[i, k].view(np.uint32).tolist()),
             int(soa.q_penalty[i, k].view(np.uint32)), int(soa.q_no_penalty[i, k].view(np.uint32))) for k in range(n)]

def run(n_games, n_slots, n_iter, kind, name, flags, spec_rows=0, arena=0, max_inline=0, n_check=64, entries=0):
    e = Engine(n_slots, n_games, n_iter, 6.6, 0.01, flags=flags, spec_rows=spec_rows, arena_blocks=arena, max_inline_sims=max_inline, eval_cache_entries=entries)
    R = e.io_rows
    io = [torch.zeros(R, 2, 6, 7, device="cuda"), torch.zeros(R, 7, device="cuda"), torch.zeros(R, device="cuda"), torch.zeros(R, device="cuda")]
    e.bind_io(*[t.data_ptr() for t in io])
    e.set_requests(list(range(n_games)), [0] * n_games, [0] * n_games)
    t = time.time()
    for i in range(10**7):
        e.eval_builtin(kind); e.step()
        if i % 64 == 63 and e.poll().n_finished == n_games: break
    got = e.fetch_results(); st = e.stats(); e.close()
    exp = oracle.self_play_parallel([(i, 0, 0) for i in range(n_check)], n_iter, 6.6, 0.01, name)
    bad = []
    for g in range(n_check):
        r = records(got, g)
        if r != exp[g]:
            k = next((k for k in range(min(len(r), len(exp[g]))) if r[k] != exp[g][k]), -1)
            bad.append((g, k))
    print(f"games={n_games} slots={n_slots} n={n_iter} {name} flags={flags} spec={spec_rows} arena={arena} inline={max_inline} entries={entries}: "
          f"{len(bad)}/{n_check} games differ, first {bad[:5]}  ticks={i+1} hits={st['cache_hits']} spec={st['spec_rows']} compactions={st['compactions']} ({time.time()-t:.1f}s)", flush=True)

C, S = L.FLAG_EVAL_CACHE, L.FLAG_SPECULATE
run(64, 64, 600, L.EVAL_HASH_FLAT, "hash_flat", 0)
run(64, 64, 600, L.EVAL_HASH_FLAT, "hash_flat", 0, arena=8 * 602)
run(64, 64, 600, L.EVAL_HASH, "hash", 0)
run(64, 64, 600, L.EVAL_HASH_FLAT, "hash_flat", C)
run(64, 128, 600, L.EVAL_HASH_FLAT, "hash_flat", C | S, spec_rows=64)
run(64, 64, 300, L.EVAL_HASH_FLAT, "hash_flat", 0)
run(64, 64, 150, L.EVAL_HASH_FLAT, "hash_flat", 0)
run(64, 64, 600, L.EVAL_UNIFORM, "uniform", 0)
run(2048, 2048, 600, L.EVAL_HASH_FLAT, "hash_flat", 0, arena=8 * 602)
run(2048, 2048, 600, L.EVAL_HASH_FLAT, "hash_flat", C, arena=8 * 602)
