"""Kernel-level check of the 3xTF32 tcgen05 GEMMs (mlp_tc.cu) against fp64 numpy: python tools/tf32x3_gemm_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ga3c_b200 import _capi
lib = _capi.load()
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
def T(x): return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dev)
def run(mode, a, b, aux, out_shape, m, k, n, act=0, splits=1, rows=32):
    out = torch.full(out_shape, float("nan"), dtype=torch.float32, device=dev)
    ta, tb = T(a), T(b); tx = T(aux) if aux is not None else None
    _capi.check(lib.ga3c_debug_tf32x3_gemm(mode, ta.data_ptr(), tb.data_ptr(), tx.data_ptr() if tx is not None else None, out.data_ptr(),
                                           m, k, n, act, splits, rows, None), "gemm")
    torch.cuda.synchronize()
    return out.cpu().numpy().astype(np.float64)
for (m, k, n) in [(128, 32, 128), (128, 256, 256), (300, 256, 100), (1000, 100, 64), (4096, 256, 256)]:
    a = rng.uniform(-1, 1, (m, k)); w = rng.uniform(-0.3, 0.3, (k, n)); bias = rng.uniform(-1, 1, n)
    a32, w32, b32 = a.astype(np.float32).astype(np.float64), w.astype(np.float32).astype(np.float64), bias.astype(np.float32).astype(np.float64)
    ref = a32 @ w32 + b32
    got = run(0, a, w, bias, (m, n), m, k, n)
    e = np.abs(got - ref); fp32 = np.abs((a.astype(np.float32) @ w.astype(np.float32) + bias.astype(np.float32)).astype(np.float64) - ref).max()
    print(f"fwd   m{m} k{k} n{n}: max err {e.max():.3e} (numpy fp32 matmul: {fp32:.3e}; max |ref| {np.abs(ref).max():.2f})", "BAD at " + str(np.unravel_index(e.argmax(), e.shape)) if not e.max() < 1e-4 else "")
    dz = rng.uniform(-1, 1, (m, n)); op = rng.uniform(0.1, 0.9, (m, k))
    dz32, op32 = dz.astype(np.float32).astype(np.float64), op.astype(np.float32).astype(np.float64)
    ref = (dz32 @ w32.T) * op32 * (1 - op32)
    got = run(1, dz, w, op, (m, k), m, k, n, act=1)
    e = np.abs(got - ref)
    print(f"dgrad m{m} k{k} n{n}: max err {e.max():.3e} (max |ref| {np.abs(ref).max():.2f})", "BAD at " + str(np.unravel_index(e.argmax(), e.shape)) if not e.max() < 1e-4 else "")
    rows = 64; splits = (m + rows - 1) // rows + 1          # one empty split at the end
    got = run(2, a, dz, None, (splits, k, n), m, k, n, splits=splits, rows=rows)
    ref = np.stack([a32[s * rows:(s + 1) * rows].T @ dz32[s * rows:(s + 1) * rows] for s in range(splits)])
    e = np.abs(got - ref)
    print(f"wgrad m{m} k{k} n{n}: max err {e.max():.3e} (max |ref| {np.abs(ref).max():.2f})", "BAD at " + str(np.unravel_index(e.argmax(), e.shape)) if not e.max() < 1e-4 else "")
