"""Prologue of conv_fwd's CTA 0 inside a pipelined run of train steps (ga3c_evt_*): kernel entry (60), barriers + TMEM
allocated (61), operands zeroed + scatter table built (62), dependency met (63), weights converted (64), and the first
frame's pipeline events.  usage: python tools/evt_conv_fwd_prologue.py"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ga3c_b200
from ga3c_b200 import _capi
tb = 1024
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
net.train_device(xs[0], yr, a)
net.predict_device(xs[1])          # its conv_fwd is the last writer of the per-warp log regions: the one that is read back
buf = (C.c_uint64 * (2 * 16384))(); cnt = C.c_int32()
_capi.check(lib.ga3c_evt_end(net._h, buf, 16384, C.byref(cnt)), "evt_end")
recs = sorted((buf[2 * i], buf[2 * i + 1] >> 32, (buf[2 * i + 1] >> 16) & 0xFFFF, buf[2 * i + 1] & 0xFFFF) for i in range(cnt.value))
t0 = recs[0][0]
names = {60: "entry", 61: "barriers + TMEM", 62: "zeroed + table", 63: "dependency met", 64: "weights converted", 15: "iss: tile rows ready", 16: "iss: tile issued", 40: "epi: conv11 tile", 52: "role done", 54: "epi: DONE passed"}
for t, w, e, arg in recs:
    if e in (60, 61, 62, 63, 64) and w in (0, 6, 8) or (e in (15, 16, 40, 44) and arg == 0 and w in (6, 8, 12)):
        print(f"{(t - t0) / 1e3:8.2f} us  warp {w:2d}  {str(names.get(e, e)):24s} {arg}")
