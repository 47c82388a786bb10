"""What would tf32 / bf16 tensor-core operands cost the fork NetworkVP MLP (BASELINE configs[3]) in accuracy?

Re-runs the forward and backward of oracle/oracle_mlp.py (fork_vp) in fp64 with the OPERANDS of the two 256-wide layers
(dense12_p 4->256 is tiny; dense13_p 256->256 and dense14_p 256->100 hold 92 % of the flops) rounded to a tensor-core input
format before every product -- forward, data gradient and weight gradient -- accumulating exactly, and reports the error
against the un-rounded run.  Decision data for a tcgen05 kind::tf32 version of mlp_fused / mlp_wgrad (experiments/ROUND2_NOTES.md).

    python experiments/tf32_mlp_error_study.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_mlp as om          # experiment only: not product code


def round_mantissa(x, bits):
    """Round-to-nearest-even to `bits` explicit mantissa bits (tf32: 10, bf16: 7), via the fp32 bit pattern."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    drop = 23 - bits
    mask = np.uint64(0xFFFFFFFF ^ ((1 << drop) - 1))
    u = (u + np.uint64((1 << (drop - 1)) - 1) + ((u >> np.uint64(drop)) & np.uint64(1))) & mask
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


def run(params, x, y_r, a, rnd, heavy=("dense13_p", "dense14_p")):
    P = {k: v.astype(np.float64) for k, v in params.items()}
    q = (lambda t: rnd(t)) if rnd else (lambda t: t)
    mm = lambda A, B, name: (q(A) @ q(B)) if name in heavy else A @ B
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))
    h = x.astype(np.float64)
    acts = [h]
    for name, _, act in om.FORK_VP_LAYERS:
        z = mm(h, P[name + "/w:0"], name) + P[name + "/b:0"]
        h = sig(z) if act == "sigmoid" else z
        acts.append(h)
    v = (h @ P["logits_v/w:0"] + P["logits_v/b:0"])[:, 0]
    ox = sig(h @ P["logits_p/out_x/w:0"] + P["logits_p/out_x/b:0"]); oy = sig(h @ P["logits_p/out_y/w:0"] + P["logits_p/out_y/b:0"])
    p = np.arctan2(oy - 0.5, ox - 0.5) / np.pi
    adv, dv, beta = y_r - v, v - y_r, 0.01
    dp = -(a * adv[:, None]) + 2.0 * beta * p
    X, Y = ox - 0.5, oy - 0.5
    r2 = X * X + Y * Y
    dzx = dp * (-Y / (np.pi * r2)) * ox * (1 - ox); dzy = dp * (X / (np.pi * r2)) * oy * (1 - oy)
    dh = dzx @ P["logits_p/out_x/w:0"].T + dzy @ P["logits_p/out_y/w:0"].T + dv[:, None] @ P["logits_v/w:0"].T
    grads = {}
    for i in range(len(om.FORK_VP_LAYERS) - 1, -1, -1):
        name, _, act = om.FORK_VP_LAYERS[i]
        out, inp = acts[i + 1], acts[i]
        dz = dh * out * (1 - out) if act == "sigmoid" else dh
        grads[name + "/w:0"] = mm(inp.T, dz, name)
        dh = mm(dz, P[name + "/w:0"].T, name)
    return p, v, grads


def main():
    rng = np.random.default_rng(0)
    params = om.init_params(rng, "fork_vp", 3, 1)
    b = 4096
    x = rng.uniform(-1, 1, (b, 3)).astype(np.float32); y_r = rng.uniform(-1, 1, b); a = rng.uniform(-1, 1, (b, 1))
    p0, v0, g0 = run(params, x, y_r, a, None)
    print(f"fork NetworkVP, S=3, A=1, B={b}, U(-0.3, 0.3) weights; operands of dense13_p / dense14_p rounded, exact accumulation")
    print("| operand format | max abs err p | max abs err v | worst gradient error / max abs gradient (tensor) |")
    print("|---|---|---|---|")
    for label, bits in (("tf32 (10-bit mantissa)", 10), ("bf16 (7-bit mantissa)", 7)):
        p1, v1, g1 = run(params, x, y_r, a, lambda t, bits=bits: round_mantissa(t, bits))
        worst = max(((np.abs(g1[k] - g0[k]).max() / np.abs(g0[k]).max(), k) for k in g0))
        print(f"| {label} | {np.abs(p1 - p0).max():.2e} | {np.abs(v1 - v0).max():.2e} | {worst[0]:.2e} ({worst[1]}) |")
    print("\nThe fp32 CUDA path is held to 2e-5 (p, v) and 2e-4 relative (gradients) against the fp64 oracle (tests/test_gpu_mlp.py).")


if __name__ == "__main__":
    main()
