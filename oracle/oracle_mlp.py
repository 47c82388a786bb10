"""numpy restatement of the two low-dimensional MLP networks of the fork (BASELINE config 4).
TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (TensorFlow 1.x is neither vendored nor installable here): pinned by finite
differences in fp64 (tests/test_oracle_mlp.py) and by torch autograd as an independent second opinion.

All paths are into /root/reference/ga3c:

  kind 'fork_vp'   NetworkVP.py:79-105, :175-192 -- x[S] -> dense 4 -> 256 -> 256 (all three LINEAR, func=None) -> 100
                   (sigmoid) -> 'dense1' 64 (sigmoid, the default func of dense_layer :194); v = dense 1 (linear);
                   p = atan2(sigmoid(out_y) - 0.5, sigmoid(out_x) - 0.5) / pi  (_create_angle_output :175-192);
                   softmax_p = log_softmax_p = logits_p (:95-96, no softmax and no log);
                   cost_p_1 = sum_a(p * a) * (R - stop_gradient(v)); cost_p_2 = -beta * sum_a(p * p)   (:97-101)
  kind 'discrate'  NetworkVP_discrate.py:52-85 -- AS WRITTEN every Config.DENSE_LAYERS entry is built from self.x (:55), so
                   only the LAST one (10 units, sigmoid) feeds the heads; the others are trainable variables without a
                   gradient (TF's minimize skips them: they keep their initial values).  Heads and loss as the conv net
                   (softmax + MIN_POLICY mix, log(max(., eps)), :66-85).
  dense_layer      NetworkVP.py:194-210: uniform(-0.3, 0.3) for w and b, output = func(x @ w + b)
  optimizer        tf.train.RMSPropOptimizer (oracle_np.rmsprop_update)

Every batch reduction is a SUM; cost_p = -(cost_p_1_agg + cost_p_2_agg); cost_all = cost_p + cost_v.
"""
from __future__ import annotations

import numpy as np

FORK_VP_LAYERS = (("dense11_p", 4, "linear"), ("dense12_p", 256, "linear"), ("dense13_p", 256, "linear"),
                  ("dense14_p", 100, "sigmoid"), ("dense1", 64, "sigmoid"))      # NetworkVP.py:79-85
DISCRATE_DENSE_LAYERS = (10, 10, 10, 10)                                          # Config.py:107


def layers_of(kind: str, dense_layers=DISCRATE_DENSE_LAYERS):
    """Live hidden layers (name, width, activation) in forward order; dense_layers = Config.DENSE_LAYERS ('discrate' only)."""
    if kind == "fork_vp":
        return FORK_VP_LAYERS
    if kind == "discrate":
        n = len(dense_layers)
        return ((f"dense1_{n}_p", dense_layers[-1], "sigmoid"),)                    # NetworkVP_discrate.py:53-56
    raise ValueError(kind)


def param_shapes(kind: str, state_dim: int, num_actions: int, dense_layers=DISCRATE_DENSE_LAYERS):
    """TF variable names in creation order -> shapes (dead 'discrate' layers included)."""
    shapes = {}
    if kind == "discrate":
        for i, w in enumerate(dense_layers[:-1]):
            shapes[f"dense1_{i + 1}_p/w:0"] = (state_dim, w)
            shapes[f"dense1_{i + 1}_p/b:0"] = (w,)
    fan = state_dim
    for name, width, _ in layers_of(kind, dense_layers):
        shapes[f"{name}/w:0"] = (fan, width)
        shapes[f"{name}/b:0"] = (width,)
        fan = width
    shapes["logits_v/w:0"] = (fan, 1)
    shapes["logits_v/b:0"] = (1,)
    if kind == "fork_vp":
        for o in ("out_x", "out_y"):
            shapes[f"logits_p/{o}/w:0"] = (fan, num_actions)
            shapes[f"logits_p/{o}/b:0"] = (num_actions,)
    else:
        shapes["logits_p/w:0"] = (fan, num_actions)
        shapes["logits_p/b:0"] = (num_actions,)
    return shapes


def dead_params(kind: str, dense_layers=DISCRATE_DENSE_LAYERS):
    """Variables that exist in the graph but receive no gradient (kept at their initial values)."""
    if kind != "discrate":
        return ()
    return tuple(f"dense1_{i + 1}_p/{t}:0" for i in range(len(dense_layers) - 1) for t in ("w", "b"))


def init_params(rng: np.random.Generator, kind: str, state_dim: int, num_actions: int, dense_layers=DISCRATE_DENSE_LAYERS):
    return {k: rng.uniform(-0.3, 0.3, size=s).astype(np.float32)
            for k, s in param_shapes(kind, state_dim, num_actions, dense_layers).items()}


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def forward(params, x, kind: str, *, dtype=np.float64, min_policy=0.0, keep=False, use_log_softmax=False,
            dense_layers=DISCRATE_DENSE_LAYERS):
    """-> p [B,A], v [B] (and the activations with keep=True)."""
    P = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    h = np.asarray(x, dtype=dtype)
    acts = [h]
    for name, _, act in layers_of(kind, dense_layers):
        z = h @ P[f"{name}/w:0"] + P[f"{name}/b:0"]
        h = _sigmoid(z) if act == "sigmoid" else z
        acts.append(h)
    v = (h @ P["logits_v/w:0"] + P["logits_v/b:0"])[:, 0]
    f = {"acts": acts, "v": v}
    if kind == "fork_vp":
        ox = _sigmoid(h @ P["logits_p/out_x/w:0"] + P["logits_p/out_x/b:0"])
        oy = _sigmoid(h @ P["logits_p/out_y/w:0"] + P["logits_p/out_y/b:0"])
        p = np.arctan2(oy - 0.5, ox - 0.5) / np.pi
        f.update(ox=ox, oy=oy)
    else:
        z = h @ P["logits_p/w:0"] + P["logits_p/b:0"]
        z = z - z.max(axis=1, keepdims=True)
        s = np.exp(z)
        s = s / s.sum(axis=1, keepdims=True)
        a_n = s.shape[1]
        p = s if use_log_softmax else (s + min_policy) / (1.0 + min_policy * a_n)      # NetworkVP_discrate.py:64-74
        f.update(s=s, lsm=np.log(s))
    f["p"] = p
    return (p, v, f) if keep else (p, v)


def loss_and_grads(params, x, y_r, a, kind: str, *, beta=0.01, log_eps=1e-6, min_policy=0.0, dtype=np.float64,
                   use_log_softmax=False, part="all", dense_layers=DISCRATE_DENSE_LAYERS):
    """-> ({cost_p_1, cost_p_2, cost_p, cost_v, cost_all}, {name: grad}) ; dead variables get no entry."""
    P = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
    y_r = np.asarray(y_r, dtype=dtype)
    a = np.asarray(a, dtype=dtype)
    p, v, f = forward(params, x, kind, dtype=dtype, min_policy=min_policy, keep=True, use_log_softmax=use_log_softmax,
                      dense_layers=dense_layers)
    adv = y_r - v                                   # stop_gradient(v) inside cost_p_1
    cost_v = 0.5 * np.sum((y_r - v) ** 2)
    dv = v - y_r
    h = f["acts"][-1]
    grads = {}
    if kind == "fork_vp":
        c1 = np.sum(np.sum(p * a, axis=1) * adv)
        c2 = np.sum(-beta * np.sum(p * p, axis=1))
        dp = -(a * adv[:, None]) + 2.0 * beta * p                      # d cost_all / d p
        X, Y = f["ox"] - 0.5, f["oy"] - 0.5
        r2 = X * X + Y * Y
        dX = dp * (-Y / (np.pi * r2))
        dY = dp * (X / (np.pi * r2))
        dzx = dX * f["ox"] * (1.0 - f["ox"])
        dzy = dY * f["oy"] * (1.0 - f["oy"])
        grads["logits_p/out_x/w:0"] = h.T @ dzx
        grads["logits_p/out_x/b:0"] = dzx.sum(axis=0)
        grads["logits_p/out_y/w:0"] = h.T @ dzy
        grads["logits_p/out_y/b:0"] = dzy.sum(axis=0)
        dh = dzx @ P["logits_p/out_x/w:0"].T + dzy @ P["logits_p/out_y/w:0"].T
    else:
        s = f["s"]
        a_n = s.shape[1]
        inv_mix = 1.0 / (1.0 + min_policy * a_n)
        sel = np.sum(p * a, axis=1)
        c1 = np.sum(np.log(np.maximum(sel, log_eps)) * adv)
        lg = np.log(np.maximum(p, log_eps))
        c2 = np.sum(-beta * np.sum(lg * p, axis=1))
        coef = np.where(sel >= log_eps, adv / np.maximum(sel, 1e-300), 0.0)
        gk = -a * coef[:, None] + beta * (lg + (p >= log_eps))          # d cost_all / d p
        hk = gk * inv_mix                                               # d / d softmax
        dz = s * (hk - np.sum(s * hk, axis=1, keepdims=True))
        if use_log_softmax:                                                 # NetworkVP_discrate.py:64-71
            lsm = f["lsm"]
            c1 = np.sum(np.sum(lsm * a, axis=1) * adv)
            c2 = np.sum(-beta * np.sum(lsm * s, axis=1))
            dz = -adv[:, None] * (a - s * a.sum(axis=1, keepdims=True)) + beta * s * (lsm - np.sum(lsm * s, axis=1, keepdims=True))
        grads["logits_p/w:0"] = h.T @ dz
        grads["logits_p/b:0"] = dz.sum(axis=0)
        dh = dz @ P["logits_p/w:0"].T
    # Config.DUAL_RMSPROP: part "p" = gradient of cost_p alone (never reaches logits_v), "v" = of cost_v alone
    if part == "p":
        dv = np.zeros_like(dv)
    elif part == "v":
        dh = np.zeros_like(dh)
        for k in list(grads):
            del grads[k]
    grads["logits_v/w:0"] = h.T @ dv[:, None]
    grads["logits_v/b:0"] = np.array([dv.sum()], dtype=dtype)
    dh = dh + dv[:, None] @ P["logits_v/w:0"].T
    layers = layers_of(kind, dense_layers)
    for i in range(len(layers) - 1, -1, -1):
        name, _, act = layers[i]
        out, inp = f["acts"][i + 1], f["acts"][i]
        dz = dh * out * (1.0 - out) if act == "sigmoid" else dh
        grads[f"{name}/w:0"] = inp.T @ dz
        grads[f"{name}/b:0"] = dz.sum(axis=0)
        dh = dz @ P[f"{name}/w:0"].T
    if part == "p":
        del grads["logits_v/w:0"], grads["logits_v/b:0"]
    cost_p = -(c1 + c2)
    losses = dict(cost_p_1=float(c1), cost_p_2=float(c2), cost_p=float(cost_p), cost_v=float(cost_v),
                  cost_all=float(cost_p + cost_v))
    return losses, grads


def train_step(params, ms, mom, x, y_r, a, kind: str, *, lr, beta=0.01, log_eps=1e-6, min_policy=0.0,
               rho=0.99, mu=0.0, eps=0.1, dtype=np.float32, grad_clip=None, dense_layers=DISCRATE_DENSE_LAYERS):
    """One opt.minimize step; dead variables and their slots are left untouched.  -> (losses, grads, params', ms', mom')"""
    from . import oracle_np as onp
    losses, grads = loss_and_grads(params, x, y_r, a, kind, beta=beta, log_eps=log_eps, min_policy=min_policy, dtype=np.float64,
                                   dense_layers=dense_layers)
    live = {k: params[k] for k in grads}
    if grad_clip is not None:              # tf.clip_by_average_norm per variable (NetworkVP.py:138-141)
        if len(grads) != len(params):
            raise ValueError("USE_GRAD_CLIP with gradient-less variables: tf.clip_by_average_norm(None, ..) raises in the reference")
        grads = {k: onp.clip_by_average_norm(g, grad_clip) for k, g in grads.items()}
    p2, ms2, mom2 = onp.rmsprop_update(live, grads, {k: ms[k] for k in grads}, {k: mom[k] for k in grads},
                                       lr=lr, rho=rho, mu=mu, eps=eps, dtype=dtype)
    out_p, out_ms, out_mom = dict(params), dict(ms), dict(mom)
    out_p.update(p2); out_ms.update(ms2); out_mom.update(mom2)
    return losses, grads, out_p, out_ms, out_mom


def train_step_dual(params, slots_p, slots_v, x, y_r, a, kind: str, *, lr, beta=0.01, log_eps=1e-6, min_policy=0.0,
                    rho=0.99, mu=0.0, eps=0.1, dtype=np.float32, grad_clip=None):
    """Config.DUAL_RMSPROP (NetworkVP.py:107-118, :143-147); semantics as oracle_np.train_step_dual: both gradients at the
    pre-call weights, w <- w - step_p - step_v, a variable a cost does not reach is skipped by that optimizer.
    grad_clip: with USE_GRAD_CLIP each optimizer's gradients go through tf.clip_by_norm per variable (NetworkVP.py:127-137)."""
    from . import oracle_np as onp
    kw = dict(beta=beta, log_eps=log_eps, min_policy=min_policy, dtype=np.float64)
    losses, gp = loss_and_grads(params, x, y_r, a, kind, part="p", **kw)
    _, gv = loss_and_grads(params, x, y_r, a, kind, part="v", **kw)
    new_p = {k: v.astype(dtype) for k, v in params.items()}
    out = []
    for g, (ms, mom) in ((gp, slots_p), (gv, slots_v)):
        if grad_clip is not None:
            g = {k: onp.clip_by_norm(v, grad_clip) for k, v in g.items()}
        p2, ms2, mom2 = onp.rmsprop_update({k: params[k] for k in g}, g, {k: ms[k] for k in g}, {k: mom[k] for k in g},
                                           lr=lr, rho=rho, mu=mu, eps=eps, dtype=dtype)
        for k in g:
            new_p[k] = new_p[k] - (params[k].astype(dtype) - p2[k])
        nms, nmom = dict(ms), dict(mom)
        nms.update(ms2); nmom.update(mom2)
        out.append((nms, nmom))
    return losses, gp, gv, new_p, out[0], out[1]
