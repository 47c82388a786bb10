"""Generates tests/golden/*.npz|json.  Run in the BUILD container only (it imports the reference's
own host-side Python from /root/reference, which does not exist on the GPU box):

    python -m oracle.gen_golden

Pinned against the REAL reference (imported, unmodified):
  returns.json      ProcessAgent._accumulate_rewards  (ProcessAgent.py:70-84) under the default
                    Config and under the three flag variants the function reads
  sampling.npz      ProcessAgent.select_action / np.random.choice (ProcessAgent.py:110-115) driven
                    by a seeded legacy RandomState, with the uniforms it consumed
  convert_data.npz  ProcessAgent.convert_data (ProcessAgent.py:86-100)
Self-pins (oracle_np output frozen so later edits cannot drift silently; the reference cannot run
its TensorFlow graph here, so these are NOT reference outputs -- parity unpinned):
  network_b4.npz    forward / losses / gradient digests / one RMSProp step at B=4, seed 12345
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF = "/root/reference/ga3c"


def import_reference_agent():
    """ProcessAgent imports gym/matplotlib/skimage transitively (EnvironmentPend ->
    PyperEnvironment -> pyper_env); stub them so the *unmodified* module imports."""
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    class _Stub(types.ModuleType):
        """Any attribute resolves to a do-nothing callable / sub-stub (simulator deps only)."""
        __path__ = []

        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return _Stub(self.__name__ + "." + item)

        def __call__(self, *a, **k):
            return None

    for name in ("gym", "gym.wrappers", "matplotlib", "matplotlib.pyplot", "matplotlib.image", "skimage",
                 "skimage.morphology", "skimage.color", "scipy.misc"):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    import ProcessAgent as PA   # noqa
    import Config as CFG        # noqa
    import Experience as EXP    # noqa
    return PA.ProcessAgent, CFG.Config, EXP.Experience


def gen_returns(ProcessAgent, Config, Experience):
    rng = np.random.default_rng(7)
    cases = []
    variants = [
        dict(DISCOUNTING=True, USE_INTERMEDIATE_REWARD=False, REWARD_CLIPPING=True),   # default
        dict(DISCOUNTING=True, USE_INTERMEDIATE_REWARD=True, REWARD_CLIPPING=True),
        dict(DISCOUNTING=False, USE_INTERMEDIATE_REWARD=False, REWARD_CLIPPING=True),
        dict(DISCOUNTING=True, USE_INTERMEDIATE_REWARD=False, REWARD_CLIPPING=False),
    ]
    saved = {k: getattr(Config, k) for k in variants[0]}
    try:
        for var in variants:
            for k, v in var.items():
                setattr(Config, k, v)
            for n in (1, 2, 3, 6, 17, 1001):
                for gamma in (0.99, 0.5):
                    rewards = [float(r) for r in rng.uniform(-3, 3, size=n)]
                    exps = [Experience(None, 0, None, r, None, False) for r in rewards]
                    terminal = rewards[-1]                       # ProcessAgent.py:148
                    out = ProcessAgent._accumulate_rewards(exps, gamma, terminal)
                    cases.append(dict(flags=var, n=n, gamma=gamma, rewards=rewards, terminal=terminal,
                                      out=[float(e.reward).hex() for e in out]))
    finally:
        for k, v in saved.items():
            setattr(Config, k, v)
    with open(os.path.join(GOLD, "returns.json"), "w") as f:
        json.dump(cases, f)
    print("returns.json", len(cases), "cases")


def gen_sampling(ProcessAgent, Config):
    assert not Config.PLAY_MODE
    rng = np.random.default_rng(11)
    n, a_n = 512, 6
    z = rng.normal(0, 2, size=(n, a_n)).astype(np.float32)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    p = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
    actions = np.arange(a_n)
    np.random.seed(4242)
    chosen = np.array([ProcessAgent.select_action(actions, p[i]) for i in range(n)], dtype=np.int32)
    np.random.seed(4242)
    u = np.array([np.random.random_sample() for _ in range(n)], dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "sampling.npz"), p=p, u=u, chosen=chosen)
    print("sampling.npz", n)


def gen_convert(ProcessAgent, Experience):
    rng = np.random.default_rng(13)
    n, s, a_n = 9, 12, 6

    class FakeSelf:       # convert_data only touches self.num_actions
        num_actions = a_n
    exps = []
    for _ in range(n):
        exps.append(Experience(rng.normal(size=s).astype(np.float32), int(rng.integers(0, a_n)),
                               None, float(rng.uniform(-1, 1)), rng.normal(size=s).astype(np.float32),
                               bool(rng.integers(0, 2))))
    x_, r_, a_, x2_, done_ = ProcessAgent.convert_data(FakeSelf(), exps)
    np.savez_compressed(os.path.join(GOLD, "convert_data.npz"),
                        states=np.array([e.state for e in exps]), actions=np.array([e.action for e in exps]),
                        rewards=np.array([e.reward for e in exps]), next_states=np.array([e.next_state for e in exps]),
                        dones=np.array([e.done for e in exps]), x_=x_, r_=r_, a_=a_, x2_=x2_, done_=done_)
    print("convert_data.npz", x_.dtype, r_.dtype, a_.dtype, done_.dtype)


def digest(g):
    g = np.asarray(g, dtype=np.float64).ravel()
    idx = np.linspace(0, g.size - 1, num=min(g.size, 64)).astype(np.int64)
    return dict(sum=float(g.sum()), l2=float(np.sqrt((g * g).sum())), amax=float(np.abs(g).max()),
                idx=idx.tolist(), val=g[idx].tolist())


def gen_network():
    from . import oracle_np as onp
    rng = np.random.default_rng(12345)                      # Config.RANDOM_SEED, Config.py:187
    params = onp.init_params(rng, 6)
    x = onp.synth_frames(rng, 4)
    y_r, a = onp.synth_targets(rng, 4)
    out = {}
    for tag, kw in (("f64", dict(dtype=np.float64)), ("bf16", dict(dtype=np.float64, quant="bf16"))):
        ms, mom = onp.rmsprop_init(params)
        losses, grads, p2, ms2, _ = onp.train_step(params, ms, mom, x, y_r, a, lr=3e-4, **kw)
        p, v = onp.forward(params, x, **kw)
        out[f"{tag}_p"], out[f"{tag}_v"] = p, v
        out[f"{tag}_losses"] = np.array([losses[k] for k in ("cost_p_1", "cost_p_2", "cost_p", "cost_v", "cost_all")])
        out[f"{tag}_digests"] = np.array(json.dumps({k: digest(g) for k, g in grads.items()}))
        out[f"{tag}_post_digests"] = np.array(json.dumps({k: digest(g) for k, g in p2.items()}))
        for k in ("conv11/b:0", "conv12/b:0", "dense1/b:0", "logits_p/w:0", "logits_v/w:0", "conv11/w:0"):
            out[f"{tag}_grad_{k}"] = grads[k]
    np.savez_compressed(os.path.join(GOLD, "network_b4.npz"), **out)
    print("network_b4.npz")


def main():
    os.makedirs(GOLD, exist_ok=True)
    ProcessAgent, Config, Experience = import_reference_agent()
    gen_returns(ProcessAgent, Config, Experience)
    gen_sampling(ProcessAgent, Config)
    gen_convert(ProcessAgent, Experience)
    gen_network()


if __name__ == "__main__":
    main()
