"""Event log of CTA 0 of the conv backward kernel for one train step (ga3c_evt_*): per-role timeline in us.
usage: python tools/evt_timeline.py [TB]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ga3c_b200
from ga3c_b200 import _capi

NAMES = {1: "ld: next quarter", 2: "ld: slot free", 20: "c12ld: next frame", 21: "c12ld: EPI12 passed",
         10: "iss: conv12 begin", 11: "iss: C12RDY passed", 12: "iss: conv12 issued", 13: "iss: conv11 begin", 14: "iss: DN1RDY passed",
         15: "iss: QFULL passed", 16: "iss: quarter issued", 17: "iss: A2RDY passed", 18: "iss: conv12 issued", 43: "epi: D drained",
         44: "epi: frame begin", 30: "b12: C12RDY passed", 40: "epi: MMA12 passed", 41: "epi: DN1FREE passed", 42: "epi: dn1 written",
         50: "prologue: at griddep_wait", 53: "prologue: dependency met", 51: "prologue: done", 52: "role done", 54: "epi: DONE passed"}
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
if os.environ.get("PREDICT"):
    net.predict_device(xs[0])
else:
    net.train_device(xs[0], yr, a)
buf = (C.c_uint64 * (2 * 16384))(); cnt = C.c_int32()
_capi.check(lib.ga3c_evt_end(net._h, buf, 16384, C.byref(cnt)), "evt_end")
recs = sorted((buf[2 * i], buf[2 * i + 1] >> 32, (buf[2 * i + 1] >> 16) & 0xFFFF, buf[2 * i + 1] & 0xFFFF) for i in range(cnt.value))
t0 = recs[0][0]
only = set(int(v) for v in os.environ.get("WARPS", "").split(",") if v)
for t, w, e, arg in recs:
    if only and w not in only: continue
    print(f"{(t - t0) / 1e3:8.2f} us  warp {w:2d}  {NAMES.get(e, e):28s} {arg}")
