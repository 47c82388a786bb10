"""Is the device-resident train loop bound by the GPU or by the host that enqueues it?  100 steps (600 launches, below the depth
of the launch queue) enqueued from Python exactly as bench.py does: host time to enqueue them vs time until the GPU is done.
usage: python tools/host_enqueue_time.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ga3c_b200

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
tb = 1024
net = ga3c_b200.Network("gpu:0", "enq", 6, max_batch=tb, seed=1)
dev = torch.device("cuda:0")
n_ring = 4
xs = [(torch.randint(0, 256, (tb, 84 * 84 * 4), device=dev, dtype=torch.int32).float() / 128 - 1).contiguous() for _ in range(n_ring)]
yr = torch.rand(tb, device=dev) * 2 - 1
a = torch.nn.functional.one_hot(torch.randint(0, 6, (tb,), device=dev), 6).float().contiguous()
stream = torch.cuda.Stream()
for i in range(20):
    net.train_device(xs[i % n_ring], yr, a, stream=stream)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(steps):
        net.train_device(xs[i % n_ring], yr, a, stream=stream)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{steps} steps: host enqueue {1e6 * (t1 - t0) / steps:7.2f} us/step, until the GPU is done {1e6 * (t2 - t0) / steps:7.2f} us/step"
          f"  ({'GPU' if (t2 - t1) > 0.1 * (t2 - t0) else 'HOST'}-bound: {1e6 * (t2 - t1):.0f} us of GPU work left when the host was done)")
