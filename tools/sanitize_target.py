"""Small jobs through every kernel path, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_target.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from c4a0_b200 import _lib as L  # noqa: E402
from c4a0_b200.engine import Engine  # noqa: E402
from c4a0_b200.nn import ConnectFourNet, FoldedNet, FusedNet, ModelConfig  # noqa: E402

for name, n_slots, n_games, flags, spec_rows, cache in (
    ("plain, refill, minimal arena", 24, 60, 0, 0, 0),
    ("cache", 40, 40, L.FLAG_EVAL_CACHE, 0, 64),
    ("cache + speculation", 40, 40, L.FLAG_EVAL_CACHE | L.FLAG_SPECULATE, 24, 0),
    ("cache + speculation, two models", 33, 50, L.FLAG_EVAL_CACHE | L.FLAG_SPECULATE, 33, 128),
):
    e = Engine(n_slots, n_games, 30, 6.6, 0.01, L.PLANES_BF16, 0, 0, 96, flags, 0, cache, spec_rows)
    R = e.io_rows
    planes = torch.zeros(R, 96, device="cuda", dtype=torch.bfloat16)
    logits = torch.zeros(R, 7, device="cuda")
    qp = torch.zeros(R, device="cuda")
    qn = torch.zeros(R, device="cuda")
    e.bind_io(planes.data_ptr(), logits.data_ptr(), qp.data_ptr(), qn.data_ptr())
    two = "two models" in name
    e.set_requests(list(range(n_games)), [1 if two else 0] * n_games, [2 if two else 0] * n_games)
    for t in range(100000):
        e.eval_builtin(L.EVAL_HASH_FLAT)
        e.step()
        if t % 16 == 15 and e.poll().n_finished == n_games:
            break
    out = e.fetch_results()
    print(name, "ticks", t + 1, "samples", int(out.n_samples.sum()), e.stats()["spec_rows"], flush=True)
    e.close()

torch.manual_seed(0)
model = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=8, n_policy_layers=3, n_value_layers=2)).cuda().eval()
for cls, dt in ((FusedNet, torch.bfloat16), (FoldedNet, torch.float32)):
    net = cls(model, dtype=dt)
    stride = getattr(net, "plane_stride", FoldedNet.IN_PAD)
    for rows in (1, 33, 4097):
        buf = torch.zeros(rows, stride, device="cuda", dtype=dt)
        out = (torch.zeros(rows, 7, device="cuda"), torch.zeros(rows, device="cuda"), torch.zeros(rows, device="cuda"))
        with torch.no_grad():
            net(buf, out=out)
        torch.cuda.synchronize()
    print(cls.__name__, "ok", flush=True)

# the shipped path: c4a0_rust.play_games -> c4a0_engine_run_net (k_step <-> k_net2 chained by programmatic dependent
# launch), minimal arenas (every re-root compacts: the work-stealing path of k_step), cache + speculation
if "--native" in sys.argv:
    import c4a0_rust  # noqa: E402
    from c4a0_b200 import selfplay  # noqa: E402

    selfplay.DEFAULTS["arena_blocks"] = 26
    m = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=8, n_policy_layers=3, n_value_layers=2)).cuda().eval()
    m = m.to(torch.bfloat16)
    reqs = [c4a0_rust.GameMetadata(i, 0, 0) for i in range(96)]
    res = c4a0_rust.play_games(reqs, 64 + 64, 24, 6.6, 0.01, m)
    print("native path:", res._run_info.ticks, "ticks,", int(res._soa.n_samples.sum()), "samples,",
          res._run_info.stats["compactions"], "compactions,", res._run_info.stats["spec_rows"], "speculative rows", flush=True)
    c4a0_rust._native.close_cached_session()
