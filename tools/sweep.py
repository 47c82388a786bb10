"""BASELINE configs[0] / configs[1] (SURVEY 8d runs 1-2): predict-only batch sweep B = 32 .. 4096 and the B = 128 predict +
train case, one GPU.  Device-resident (CUDA events on the launch stream, inputs rotated through a ring larger than L2) and
end to end through Network.predict_p_and_v / Network.train on pinned host buffers.  Prints a markdown table.
usage: python tools/sweep.py > profiles/<round>_predict_sweep.md"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ga3c_b200

S, A, L2 = 84 * 84 * 4, 6, 126 * 2 ** 20
dev = torch.device("cuda:0")
net = ga3c_b200.Network("gpu:0", "sweep", A, max_batch=4096, seed=12345)
st = torch.cuda.current_stream(dev)
rng = np.random.default_rng(12345)
peak = 6546.6
try:
    import json
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def frames(b):
    k = rng.integers(0, 256, size=(b, S), dtype=np.uint8)
    return (k.astype(np.float32) / np.float32(128.0) - np.float32(1.0))


def dev_time(fn, steps):
    for i in range(5):
        fn(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for i in range(steps):
        fn(i)
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3          # us


def host_time(fn, steps):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e6


print("# Predict-only batch sweep and the B = 128 case, 1 B200 (tools/sweep.py)\n")
print("Device-resident: CUDA events on the launch stream, 200 calls, inputs rotated through a ring of batches larger than the "
      "126 MB L2.  e2e: Network.predict_p_and_v / Network.train on pinned host numpy (H2D + D2H inside), 30 calls.  Roofline = "
      f"112,896 B/frame (fp32 frame read once) against {peak:.0f} GB/s measured HBM peak.\n")
print("| B | predict us | predictions/s | frac of HBM roofline | e2e predict us | e2e predictions/s |")
print("|---|---|---|---|---|---|")
for b in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    n_ring = max(2, int(1.5 * L2) // (b * S * 4) + 1)
    base = torch.from_numpy(frames(min(b * 4, 4096))).to(dev)
    ring = [torch.roll(base, shifts=i, dims=0)[:b].contiguous() for i in range(min(n_ring, 96))]
    p = torch.empty((b, A), device=dev); v = torch.empty((b,), device=dev)
    us = dev_time(lambda i: net.predict_device(ring[i % len(ring)], p, v, stream=st), 200)
    hx = torch.from_numpy(frames(b)).pin_memory()
    us_e = host_time(lambda i: net.predict_p_and_v(hx.numpy()), 30)
    print(f"| {b} | {us:.1f} | {b / us * 1e6 / 1e6:.2f} M | {b * S * 4 / (us * 1e-6) / 1e9 / peak:.3f} | {us_e:.0f} | {b / us_e * 1e6 / 1e3:.0f} k |")
    del ring, base
print("\n| B | train step us | training frames/s | e2e train us | e2e frames/s |")
print("|---|---|---|---|---|")
for b in (128, 1024):
    n_ring = max(2, int(1.5 * L2) // (b * S * 4) + 1)
    base = torch.from_numpy(frames(min(b * 4, 4096))).to(dev)
    ring = [torch.roll(base, shifts=i, dims=0)[:b].contiguous() for i in range(min(n_ring, 96))]
    yr = torch.rand(b, device=dev) * 2 - 1
    a = torch.nn.functional.one_hot(torch.randint(0, A, (b,), device=dev), A).float().contiguous()
    us = dev_time(lambda i: net.train_device(ring[i % len(ring)], yr, a, stream=st), 200)
    hx = torch.from_numpy(frames(b)).pin_memory(); hyr = yr.cpu().numpy(); ha = a.cpu().numpy()
    us_e = host_time(lambda i: net.train(hx.numpy(), hyr, ha, None, None, 0), 30)
    print(f"| {b} | {us:.1f} | {b / us:.2f} M | {us_e:.0f} | {b / us_e * 1e3:.0f} k |")
    del ring, base
