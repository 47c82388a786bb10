"""A few training steps of the fork NetworkVP MLP at B = 65,536 (BASELINE configs[3]) -- the command profiled with ncu.
usage: python tools/mlp_step.py [steps] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ga3c_b200 import mlp_network
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
b = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
rng = np.random.default_rng(0)
net = mlp_network.NetworkVP("gpu:0", "mlpstep", 1, 3, max_batch=b)
dev = torch.device("cuda", 0)
x = torch.from_numpy(rng.uniform(-1, 1, (b, 3)).astype(np.float32)).to(dev)
yr = torch.from_numpy(rng.uniform(-1, 1, b).astype(np.float32)).to(dev)
a = torch.from_numpy(rng.uniform(-1, 1, (b, 1)).astype(np.float32)).to(dev)
for _ in range(steps):
    net.train_device(x, yr, a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    net.train_device(x, yr, a)
e1.record()
torch.cuda.synchronize()
print(f"B={b}: {e0.elapsed_time(e1) / steps * 1e3:.1f} us/step, {b * steps / e0.elapsed_time(e1) / 1e3:.1f} M samples/s")
