"""Where does the end-to-end step spend its host time beyond the device-timed search?  cProfile of one
bench step (c4a0_rust.play_games with host requests in, host samples out) after two warm-up steps."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import c4a0_rust  # noqa: E402
from c4a0_b200.nn import ConnectFourNet, default_config  # noqa: E402
from c4a0_b200.selfplay import DeviceEvaluator  # noqa: E402

G, sims = 16384, 600
torch.manual_seed(1337)
model = ConnectFourNet(default_config()).cuda().eval()
state = {"ev": None}


def one_step():
    ev = DeviceEvaluator.from_model(model, torch.bfloat16, reuse=state["ev"])
    state["ev"] = ev
    reqs = [c4a0_rust.GameMetadata(i, 0, 0) for i in range(G)]
    res = c4a0_rust.play_games(reqs, G + 8192, sims, 6.6, 0.01, ev)
    return res


for _ in range(2):
    one_step()
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
res = one_step()
pr.disable()
dt = time.perf_counter() - t0
print(f"wall {dt * 1e3:.1f} ms, device-timed search {res._run_info.device_s * 1e3:.1f} ms, host remainder {(dt - res._run_info.device_s) * 1e3:.1f} ms")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
