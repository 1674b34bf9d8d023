"""A few forward passes of the native network kernel at one batch size (target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from c4a0_b200.native_net import NativeEvaluator
from c4a0_b200.nn import ConnectFourNet, ModelConfig

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5120
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
torch.manual_seed(1337)
dev = torch.device("cuda", 0)
model = ConnectFourNet(ModelConfig(n_residual_blocks=1, conv_filter_size=32, n_policy_layers=4, n_value_layers=2)).to(dev).eval()
net = NativeEvaluator(model).instantiate(max(rows, 256))
net(torch.zeros(rows, 84, device=dev))
net.buffer(0)[:rows, 1344:1344 + 84] = (torch.rand(rows, 84, device=dev) < 0.25).to(torch.bfloat16)
for _ in range(n):
    net.forward(rows)
torch.cuda.synchronize()
print("done")
