"""Pins for the config-4 MLP oracle (oracle/oracle_mlp.py): finite differences in fp64 and torch autograd as an
independent second opinion (the reference's TensorFlow is not installable: parity unpinned)."""
import numpy as np
import pytest
import torch

from oracle import oracle_mlp as om

CASES = [("fork_vp", 3, 1), ("fork_vp", 4, 2), ("discrate", 4, 2)]


def _case(kind, s, a, b=9, seed=5):
    rng = np.random.default_rng(seed)
    params = om.init_params(rng, kind, s, a)
    x = rng.uniform(-1, 1, size=(b, s)).astype(np.float32)
    y_r = rng.uniform(-1, 1, size=b).astype(np.float32)
    if kind == "fork_vp":
        act = rng.uniform(-1, 1, size=(b, a)).astype(np.float32)          # continuous action (ProcessAgent.py:92-93)
    else:
        act = np.eye(a, dtype=np.float32)[rng.integers(0, a, size=b)]
    return params, x, y_r, act


@pytest.mark.parametrize("kind,s,a", CASES)
def test_analytic_gradients_match_finite_differences(kind, s, a):
    params, x, y_r, act = _case(kind, s, a)
    p64 = {k: v.astype(np.float64) for k, v in params.items()}
    _, grads = om.loss_and_grads(p64, x, y_r, act, kind)
    assert set(grads) == set(params) - set(om.dead_params(kind))
    rng = np.random.default_rng(1)
    for name, g in grads.items():
        for _ in range(4):
            idx = tuple(rng.integers(0, d) for d in g.shape)
            h = 1e-6
            pp = {k: v.copy() for k, v in p64.items()}; pp[name][idx] += h
            pm = {k: v.copy() for k, v in p64.items()}; pm[name][idx] -= h
            # cost_p_1 uses stop_gradient(v): differentiate with the advantage frozen at the unperturbed v
            fd = (_frozen_cost(pp, p64, x, y_r, act, kind) - _frozen_cost(pm, p64, x, y_r, act, kind)) / (2 * h)
            assert abs(fd - g[idx]) <= 1e-5 * max(1.0, abs(fd)), (name, idx, fd, g[idx])


def _frozen_cost(p, p_ref, x, y_r, act, kind, beta=0.01, log_eps=1e-6):
    pr, v = om.forward(p, x, kind)
    _, v0 = om.forward(p_ref, x, kind)
    adv = y_r - v0
    cost_v = 0.5 * np.sum((y_r - v) ** 2)
    if kind == "fork_vp":
        c1 = np.sum(np.sum(pr * act, axis=1) * adv)
        c2 = np.sum(-beta * np.sum(pr * pr, axis=1))
    else:
        c1 = np.sum(np.log(np.maximum(np.sum(pr * act, axis=1), log_eps)) * adv)
        c2 = np.sum(-beta * np.sum(np.log(np.maximum(pr, log_eps)) * pr, axis=1))
    return -(c1 + c2) + cost_v


@pytest.mark.parametrize("kind,s,a", CASES)
def test_torch_autograd_second_opinion(kind, s, a):
    params, x, y_r, act = _case(kind, s, a, b=17, seed=9)
    losses, grads = om.loss_and_grads(params, x, y_r, act, kind)
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params.items()}
    h = torch.tensor(x, dtype=torch.float64)
    for name, _, fn in om.layers_of(kind):
        h = h @ P[f"{name}/w:0"] + P[f"{name}/b:0"]
        if fn == "sigmoid":
            h = torch.sigmoid(h)
    v = (h @ P["logits_v/w:0"] + P["logits_v/b:0"])[:, 0]
    R, A = torch.tensor(y_r, dtype=torch.float64), torch.tensor(act, dtype=torch.float64)
    adv = R - v.detach()
    if kind == "fork_vp":
        ox = torch.sigmoid(h @ P["logits_p/out_x/w:0"] + P["logits_p/out_x/b:0"])
        oy = torch.sigmoid(h @ P["logits_p/out_y/w:0"] + P["logits_p/out_y/b:0"])
        p = torch.atan2(oy - 0.5, ox - 0.5) / np.pi
        c1 = ((p * A).sum(1) * adv).sum()
        c2 = (-0.01 * (p * p).sum(1)).sum()
    else:
        p = torch.softmax(h @ P["logits_p/w:0"] + P["logits_p/b:0"], dim=1)
        c1 = (torch.log(torch.clamp((p * A).sum(1), min=1e-6)) * adv).sum()
        c2 = (-0.01 * (torch.log(torch.clamp(p, min=1e-6)) * p).sum(1)).sum()
    cost_v = 0.5 * ((R - v) ** 2).sum()
    cost_all = -(c1 + c2) + cost_v
    cost_all.backward()
    assert abs(float(cost_all) - losses["cost_all"]) <= 1e-9 * max(1.0, abs(losses["cost_all"]))
    for name, g in grads.items():
        tg = P[name].grad.numpy()
        assert np.abs(tg - g).max() <= 1e-9 * max(1.0, np.abs(g).max()), name
    for name in om.dead_params(kind):
        assert P[name].grad is None


def test_discrate_dead_layers_keep_their_values():
    params, x, y_r, act = _case("discrate", 4, 2)
    from oracle import oracle_np as onp
    ms, mom = onp.rmsprop_init(params)
    _, _, p2, ms2, _ = om.train_step(params, ms, mom, x, y_r, act, "discrate", lr=3e-4)
    for name in om.dead_params("discrate"):
        assert np.array_equal(p2[name], params[name]) and np.array_equal(ms2[name], ms[name])
    assert not np.array_equal(p2["dense1_4_p/w:0"], params["dense1_4_p/w:0"])


def test_discrate_log_softmax_branch_matches_torch():
    """Config.USE_LOG_SOFTMAX (NetworkVP_discrate.py:64-71) against torch autograd."""
    kind, s, a = "discrate", 4, 2
    params, x, y_r, act = _case(kind, s, a, b=11, seed=8)
    losses, grads = om.loss_and_grads(params, x, y_r, act, kind, beta=0.02, use_log_softmax=True)
    tp = {k: torch.tensor(v.astype(np.float64), requires_grad=True) for k, v in params.items()}
    tx, tyr, ta = (torch.tensor(np.asarray(t, dtype=np.float64)) for t in (x, y_r, act))
    h = torch.sigmoid(tx @ tp["dense1_4_p/w:0"] + tp["dense1_4_p/b:0"])
    v = (h @ tp["logits_v/w:0"] + tp["logits_v/b:0"])[:, 0]
    z = h @ tp["logits_p/w:0"] + tp["logits_p/b:0"]
    lsm, sm = torch.log_softmax(z, 1), torch.softmax(z, 1)
    c1 = ((lsm * ta).sum(1) * (tyr - v.detach())).sum()
    c2 = (-0.02 * (lsm * sm).sum(1)).sum()
    total = -(c1 + c2) + 0.5 * ((tyr - v) ** 2).sum()
    total.backward()
    assert abs(float(total) - losses["cost_all"]) <= 1e-10 * max(1.0, abs(losses["cost_all"]))
    for k, g in grads.items():
        assert np.abs(tp[k].grad.numpy() - g).max() <= 1e-11, k


def _golden_case(g, kind):
    s_, a_ = (3, 1) if kind == "fork_vp" else (4, 2)
    params = om.init_params(np.random.default_rng(2024), kind, s_, a_)          # as oracle/gen_golden.py: gen_mlp drew them
    grads = {k[len(kind) + 6:]: g[k] for k in g.files if k.startswith(kind + "_grad_")}
    return params, g[kind + "_x"], g[kind + "_yr"], g[kind + "_a"], grads


@pytest.mark.parametrize("kind", ["fork_vp", "discrate"])
def test_oracle_matches_golden_fixture(kind, golden_dir):
    """tests/golden/mlp_b6.npz: a torch-autograd restatement of the reference graph lines (oracle/gen_golden.py: gen_mlp)."""
    import os
    g = np.load(os.path.join(golden_dir, "mlp_b6.npz"))
    params, x, y_r, act, grads_ref = _golden_case(g, kind)
    p, v = om.forward(params, x, kind)
    assert np.abs(p - g[kind + "_p"]).max() <= 1e-12 and np.abs(v - g[kind + "_v"]).max() <= 1e-12
    losses, grads = om.loss_and_grads(params, x, y_r, act, kind, beta=0.01)
    got = np.array([losses[k] for k in ("cost_p_1", "cost_p_2", "cost_p", "cost_v", "cost_all")])
    assert np.allclose(got, g[kind + "_losses"], rtol=1e-11, atol=1e-13)
    assert set(grads) == set(grads_ref)                      # gradient-less variables have no entry on either side
    for k in grads:
        assert np.abs(grads[k] - grads_ref[k]).max() <= 1e-6 * max(1.0, np.abs(grads_ref[k]).max()), k   # fixture is float32


@pytest.mark.parametrize("kind,s,a", CASES)
def test_dual_rmsprop_gradients_add_up(kind, s, a):
    """Config.DUAL_RMSPROP: grad(cost_p) + grad(cost_v) = the (pinned) grad(cost_all); each cost skips the other head."""
    params, x, y_r, act = _case(kind, s, a)
    _, g_all = om.loss_and_grads(params, x, y_r, act, kind)
    _, g_p = om.loss_and_grads(params, x, y_r, act, kind, part="p")
    _, g_v = om.loss_and_grads(params, x, y_r, act, kind, part="v")
    assert set(g_all) - set(g_p) == {"logits_v/w:0", "logits_v/b:0"}
    assert set(g_all) - set(g_v) == {k for k in g_all if k.startswith("logits_p")}
    for k in g_all:
        tot = g_p.get(k, 0.0) + g_v.get(k, 0.0)
        assert np.abs(tot - g_all[k]).max() <= 1e-12 * max(1.0, np.abs(g_all[k]).max()), k
