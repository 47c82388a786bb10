"""Tiny driver for ncu captures of the config-4 MLP kernels: a few predict + train steps at B (default 65536)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ga3c_b200 import mlp_network
b = int(os.environ.get("B", 65536)); reps = int(os.environ.get("REPS", 2))
kind = os.environ.get("KIND", "fork_vp")
dev = torch.device("cuda:0")
net = (mlp_network.NetworkVP("gpu:0", "prof", 1, 3, max_batch=b, seed=1) if kind == "fork_vp"
       else mlp_network.NetworkVP_discrate("gpu:0", "prof", 2, 4, max_batch=b, seed=1))
s, a = net.state_dim, net.num_actions
x = torch.rand(b, s, device=dev) * 2 - 1
yr = torch.rand(b, device=dev) * 2 - 1
act = torch.rand(b, a, device=dev) * 2 - 1 if kind == "fork_vp" else torch.nn.functional.one_hot(torch.randint(0, a, (b,), device=dev), a).float().contiguous()
for _ in range(reps):
    net.predict_device(x)
for _ in range(reps):
    net.train_device(x, yr, act)
torch.cuda.synchronize()
print("ok", net.launch_count())
if os.environ.get("TIME"):
    for name, fn in (("predict", lambda: net.predict_device(x)), ("train", lambda: net.train_device(x, yr, act))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{kind} B={b} {name}: {ms * 1e3:.1f} us  {b / ms / 1e3:.2f} M/s")
    net.kernel_timing(200)
    for _ in range(20): net.train_device(x, yr, act)
    print({k: round(t / c * 1e3, 1) for k, (t, c) in net.kernel_times().items()})
