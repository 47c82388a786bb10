"""Event log of the first tile of each dense1 GEMM (forward, data gradient, weight gradient) inside one pipelined train step.
Needs a debug build:  GA3C_NVCC_EXTRA=-DGA3C_DENSE_EVT python -c "from ga3c_b200.build import build_library as b; b(force=True)"
(rebuild without the flag afterwards); run with GA3C_EVT_DENSE=1.   usage: GA3C_EVT_DENSE=1 python tools/evt_dense.py [TB]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GA3C_EVT_DENSE", "1")
import torch
import ga3c_b200
from ga3c_b200 import _capi

NAMES = {0: "dependency met", 1: "producer: slot free", 2: "producer: boxes issued", 3: "mma: stage landed", 4: "mma: issued + commit",
         5: "epi: prefetch issued", 6: "epi: accumulator complete", 7: "epi: chunk stored", 8: "tile done"}
ROLE = {0: "fwd producer", 1: "fwd mma", 2: "fwd epi0", 3: "fwd epi1", 4: "fwd epi2", 5: "fwd epi3",
        6: "dgrad producer", 7: "dgrad mma", 8: "dgrad epi0", 9: "dgrad epi1", 10: "dgrad epi2", 11: "dgrad epi3",
        12: "wgrad producer", 13: "wgrad mma", 14: "wgrad epi0", 15: "wgrad epi1"}
tb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = ga3c_b200.Network("gpu:0", "evt", 6, max_batch=tb, seed=1)
dev = torch.device("cuda:0")
xs = [(torch.randint(0, 256, (tb, 84 * 84 * 4), device=dev, dtype=torch.int32).float() / 128 - 1).contiguous() for _ in range(3)]
yr = torch.rand(tb, device=dev) * 2 - 1
a = torch.nn.functional.one_hot(torch.randint(0, 6, (tb,), device=dev), 6).float().contiguous()
for i in range(10):
    net.train_device(xs[i % 3], yr, a)
torch.cuda.synchronize()
lib = _capi.load()
_capi.check(lib.ga3c_evt_begin(net._h), "evt_begin")
net.train_device(xs[1], yr, a)
net.train_device(xs[2], yr, a)          # the log keeps the LAST step's records (cursors restart per kernel): a steady-state step
buf = (C.c_uint64 * (2 * 16384))(); cnt = C.c_int32()
_capi.check(lib.ga3c_evt_end(net._h, buf, 16384, C.byref(cnt)), "evt_end")
recs = sorted((buf[2 * i], buf[2 * i + 1] >> 32, (buf[2 * i + 1] >> 16) & 0xFFFF, buf[2 * i + 1] & 0xFFFF) for i in range(cnt.value))
if not recs:
    sys.exit("no records: not a -DGA3C_DENSE_EVT build")
for lo, hi, name in ((0, 6, "dense_fwd tile (0,0,0)"), (6, 12, "dense_bwd: data-gradient tile 0"), (12, 16, "dense_bwd: weight-gradient tile 0")):
    sub = [r for r in recs if lo <= r[1] < hi]
    if not sub:
        continue
    t0 = sub[0][0]
    print(f"### {name}  (t = 0 at its first record)")
    for t, w, e, arg in sub:
        print(f"{(t - t0) / 1e3:8.2f} us  {ROLE.get(w, w):16s} {NAMES.get(e, e):28s} {arg}")
