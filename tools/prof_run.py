"""Tiny driver for ncu captures: a few predict (B=4096) and train (B=1024) steps on device-resident data."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ga3c_b200
pb = int(os.environ.get("PB", 4096)); tb = int(os.environ.get("TB", 1024)); reps = int(os.environ.get("REPS", 3))
net = ga3c_b200.Network("gpu:0", "prof", 6, max_batch=max(pb, tb), seed=1)
dev = torch.device("cuda:0")
x = (torch.randint(0, 256, (max(pb, tb), 84 * 84 * 4), device=dev, dtype=torch.int32).float() / 128 - 1).contiguous()
yr = torch.rand(tb, device=dev) * 2 - 1
a = torch.nn.functional.one_hot(torch.randint(0, 6, (tb,), device=dev), 6).float().contiguous()
for _ in range(reps):
    net.predict_device(x[:pb])
for _ in range(reps):
    net.train_device(x[:tb], yr, a)
torch.cuda.synchronize()
print("ok", net.launch_count())
