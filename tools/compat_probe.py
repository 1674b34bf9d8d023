"""Throughput of the numpy-callback compatibility path: what an UNMODIFIED `main.py train` takes
(src/c4a0/training.py:180-189 hands `model.forward_numpy` to c4a0_rust.play_games).  Every tick the live
rows go to the host, through the callback (fp32 module on cuda:0) and back.  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import c4a0_rust  # noqa: E402
from c4a0_b200.nn import ConnectFourNet, default_config  # noqa: E402

games = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 600
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
torch.manual_seed(1337)
model = ConnectFourNet(default_config()).cuda().eval()
reqs = [c4a0_rust.GameMetadata(i, 0, 0) for i in range(games)]
c4a0_rust.play_games(reqs[:64], batch, 32, 6.6, 0.01, lambda mid, x: model.forward_numpy(x))  # warm-up
c4a0_rust._native.close_cached_session()
t0 = time.perf_counter()
res = c4a0_rust.play_games(reqs, batch, sims, 6.6, 0.01, lambda mid, x: model.forward_numpy(x))
dt = time.perf_counter() - t0
st = res._run_info.stats
print(json.dumps({"path": "numpy callback (compat)", "games": games, "sims_per_move": sims, "max_nn_batch_size": batch,
                  "seconds": dt, "positions": int(st["samples"]), "positions_per_s": st["samples"] / dt,
                  "sims_per_s": st["sims"] / dt, "ticks": res._run_info.ticks, "ms_per_tick": 1e3 * dt / max(1, res._run_info.ticks)}))
# the same job on the device fast path (module handed to play_games: fp32 parameters -> fp32 PyTorch network on the device)
c4a0_rust._native.close_cached_session()
for dtype in (torch.float32, torch.bfloat16):
    m = ConnectFourNet(default_config())
    m.load_state_dict(model.state_dict())
    m = m.to(device="cuda", dtype=dtype).eval()
    c4a0_rust.play_games(reqs, batch, sims, 6.6, 0.01, m)  # warm-up: engine + graphs
    t0 = time.perf_counter()
    res = c4a0_rust.play_games(reqs, batch, sims, 6.6, 0.01, m)
    dt = time.perf_counter() - t0
    st = res._run_info.stats
    print(json.dumps({"path": f"device evaluator ({dtype})", "games": games, "sims_per_move": sims, "max_nn_batch_size": batch,
                      "seconds": dt, "positions": int(st["samples"]), "positions_per_s": st["samples"] / dt,
                      "sims_per_s": st["sims"] / dt, "ticks": res._run_info.ticks}))
    c4a0_rust._native.close_cached_session()
