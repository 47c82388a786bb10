"""Times the 3xTF32 GEMM launches of mlp_tc.cu at the fork NetworkVP's shapes (B = 65,536): python tools/tf32x3_gemm_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ga3c_b200 import _capi
lib = _capi.load()
dev = torch.device("cuda", 0)
m = 65536
def T(*shape): return (torch.rand(*shape, device=dev) - 0.5).contiguous()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
st = torch.cuda.current_stream().cuda_stream
for (k, n) in [(256, 256), (256, 100), (100, 64)]:
    a, w, bias, out = T(m, k), T(k, n), T(n), torch.empty(m, n, device=dev)
    dz, op, dprev = T(m, n), torch.rand(m, k, device=dev), torch.empty(m, k, device=dev)
    splits, rows = 74, 896
    part = torch.empty(splits, k, n, device=dev)
    f = lambda: _capi.check(lib.ga3c_debug_tf32x3_gemm(0, a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), m, k, n, 0, 1, 32, st), "f")
    d = lambda: _capi.check(lib.ga3c_debug_tf32x3_gemm(1, dz.data_ptr(), w.data_ptr(), op.data_ptr(), dprev.data_ptr(), m, k, n, 1, 1, 32, st), "d")
    g = lambda: _capi.check(lib.ga3c_debug_tf32x3_gemm(2, a.data_ptr(), dz.data_ptr(), None, part.data_ptr(), m, k, n, 0, splits, rows, st), "g")
    flop = 2.0 * m * k * n
    tf, td, tg = timeit(f), timeit(d), timeit(g)
    print(f"k{k} n{n}: fwd {tf:6.1f} us ({flop/tf/1e6:6.1f} TF fp32-equiv)  dgrad {td:6.1f} us ({flop/td/1e6:6.1f})  wgrad {tg:6.1f} us ({flop/tg/1e6:6.1f})", flush=True)
