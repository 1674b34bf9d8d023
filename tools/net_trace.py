"""Event trace of one CTA of the network kernel (cycle deltas per role)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from c4a0_b200.native_net import NativeEvaluator
from c4a0_b200.nn import ConnectFourNet, ModelConfig
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if len(sys.argv) > 3: os.environ["C4A0_NET_VARIANT"] = sys.argv[3]
torch.manual_seed(1337)
dev = torch.device("cuda", 0)
model = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=32, n_policy_layers=4, n_value_layers=2)).to(dev).eval()
net = NativeEvaluator(model).instantiate(max(rows, 256))
net(torch.zeros(max(rows, 1), 84, device=dev))
for _ in range(3): net.forward(rows)
torch.cuda.synchronize()
tr = net.debug_trace(rows, cta)
t0 = min(r[0][1] for r in tr if r)
names = ["producer", "mma", "epilogue"]
for name, role in zip(names, tr):
    print(f"--- {name}: {len(role)} events")
    prev = t0
    line = []
    for tag, t in role[:400]:
        line.append(f"{tag}@{t - t0}(+{t - prev})")
        prev = t
    for i in range(0, len(line), 8):
        print("  " + "  ".join(line[i:i + 8]))
