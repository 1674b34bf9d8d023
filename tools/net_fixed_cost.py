import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from c4a0_b200.native_net import NativeEvaluator
from c4a0_b200.nn import ConnectFourNet, ModelConfig
torch.manual_seed(1337)
dev = torch.device("cuda", 0)
model = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=32, n_policy_layers=4, n_value_layers=2)).to(dev).eval()
ev = NativeEvaluator(model)
for cfg in ("pair", "pair:1", "pair:3", "one"):
    os.environ.pop("C4A0_NET_ONE_CTA", None); os.environ.pop("C4A0_NET_VARIANT", None)
    if cfg == "one": os.environ["C4A0_NET_ONE_CTA"] = "1"
    elif ":" in cfg: os.environ["C4A0_NET_VARIANT"] = cfg.split(":")[1]
    net = ev.instantiate(4096)
    net(torch.zeros(8, 84, device=dev))
    for rows in (0, 1, 32, 33, 128, 129, 256, 512, 1024, 2048, 4096):
        for _ in range(3): net.forward(rows)
        torch.cuda.synchronize()
        ts = sorted(net.forward_timed(rows) for _ in range(40))
        print(f"{cfg:7s} rows {rows:5d}: median {1e3*ts[20]:7.1f} us  min {1e3*ts[0]:7.1f} us", flush=True)
    net.close()
# an empty kernel for reference: launch + event overhead
x = torch.zeros(1, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(40):
    e0.record(); x.add_(1); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
print("tiny torch kernel between events:", 1e3 * sorted(ts)[20], "us")
