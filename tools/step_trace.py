"""Device-side timeline of one pipelined A3C train step (and one predict call): every CTA stamps %globaltimer when it
is launched, when its dependency wait returns and when it ends (ga3c_trace_*).  Unlike per-kernel events or ncu this
does not serialise the programmatic-dependent-launch chain.  usage: python tools/step_trace.py [TB] [PB] > profiles/x.md"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ga3c_b200

tb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pb = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
net = ga3c_b200.Network("gpu:0", "trace", 6, max_batch=max(pb, tb), seed=1)
dev = torch.device("cuda:0")
n_ring = 3
xs = [(torch.randint(0, 256, (max(pb, tb), 84 * 84 * 4), device=dev, dtype=torch.int32).float() / 128 - 1).contiguous() for _ in range(n_ring)]
yr = torch.rand(tb, device=dev) * 2 - 1
a = torch.nn.functional.one_hot(torch.randint(0, 6, (tb,), device=dev), 6).float().contiguous()


def show(title, rows):
    print(f"\n### {title}\n")
    print("| kernel | first CTA launched | last launched | first started | last started | first ended | last ended | busy (first start -> last end) |")
    print("|---|---|---|---|---|---|---|---|")
    for k, v in sorted(rows.items(), key=lambda kv: kv[1][2]):
        print(f"| `{k}` | " + " | ".join(f"{t:.1f}" for t in v) + f" | {v[5] - v[2]:.1f} |")
    end = max(v[5] for v in rows.values())
    print(f"\nspan: {end:.1f} us (all times in us from the first stamp)")


for i in range(20):
    net.train_device(xs[i % n_ring][:tb], yr, a)
for reps in (1, 2):
    torch.cuda.synchronize()
    net.trace_begin()
    for i in range(reps):
        net.train_device(xs[i % n_ring][:tb], yr, a)
    show(f"train step, B = {tb}" + ("" if reps == 1 else f" -- {reps} steps back to back (min / max over both)"), net.trace_end())
for i in range(5):
    net.predict_device(xs[i % n_ring][:pb])
torch.cuda.synchronize()
net.trace_begin()
net.predict_device(xs[0][:pb])
show(f"predict, B = {pb}", net.trace_end())
