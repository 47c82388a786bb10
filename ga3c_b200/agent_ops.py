"""Device versions of the two per-agent computations on the path.

    accumulate_rewards   ProcessAgent._accumulate_rewards  (ProcessAgent.py:70-84)
    select_actions       ProcessAgent.select_action        (ProcessAgent.py:110-115), given the uniforms

Both are fp64 on the device with the reference's operation order, so results are bit-identical to
the Python/numpy reference.  Many agents' rollouts are processed in one launch (ragged segments).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi
from .config import Config as _DefaultConfig


def _flags(cfg, nstep=False):
    f = 0
    if cfg.DISCOUNTING:
        f |= _capi.RET_DISCOUNTING
    if cfg.USE_INTERMEDIATE_REWARD:
        f |= _capi.RET_INTERMEDIATE
    if cfg.REWARD_CLIPPING:
        f |= _capi.RET_CLIPPING
    if nstep:
        f |= _capi.RET_NSTEP
    return f


def accumulate_rewards(reward_segments, discount, terminal_rewards, *, config=None, nstep=False, device="cuda:0"):
    """reward_segments: list of 1-D float sequences (one per agent rollout, any lengths incl. 0);
    terminal_rewards: one float per segment (ProcessAgent.py:148 passes the last reward).
    Returns a list of float64 arrays: the `.reward` fields after the reference's in-place update.
    With nstep=True: upstream R_t = clip(r_t) + gamma R_{t+1} seeded by the terminal value."""
    cfg = config or _DefaultConfig
    lib = _capi.load()
    lens = np.array([len(s) for s in reward_segments], dtype=np.int64)
    offs = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    total = int(offs[-1])
    if len(lens) == 0:
        return []
    flat = np.concatenate([np.asarray(s, dtype=np.float64).ravel() for s in reward_segments]) if total else np.zeros(0)
    dev = torch.device(device)
    with torch.cuda.device(dev):
        d_r = torch.from_numpy(flat).to(dev)
        d_o = torch.from_numpy(offs).to(dev)
        d_t = torch.from_numpy(np.asarray(terminal_rewards, dtype=np.float64)).to(dev)
        d_out = torch.empty(max(total, 1), dtype=torch.float64, device=dev)
        st = torch.cuda.current_stream(dev)
        _capi.check(lib.ga3c_returns(d_r.data_ptr(), d_o.data_ptr(), len(lens), d_t.data_ptr(), float(discount),
                                     _flags(cfg, nstep), float(cfg.REWARD_MIN), float(cfg.REWARD_MAX),
                                     d_out.data_ptr(), st.cuda_stream), "ga3c_returns")
        out = d_out.cpu().numpy()
    return [out[offs[i]:offs[i + 1]].copy() for i in range(len(lens))]


def select_actions(p, u, *, device="cuda:0"):
    """p: float32 [B, A] policies; u: float64 [B] uniforms (what np.random.random_sample() returned).
    Returns int32 [B] -- identical to np.random.choice(arange(A), p=p[i]) consuming u[i]."""
    lib = _capi.load()
    p = np.ascontiguousarray(p, dtype=np.float32)
    u = np.ascontiguousarray(u, dtype=np.float64)
    b, a = p.shape
    if b == 0:
        return np.zeros(0, dtype=np.int32)
    dev = torch.device(device)
    with torch.cuda.device(dev):
        d_p = torch.from_numpy(p).to(dev)
        d_u = torch.from_numpy(u).to(dev)
        d_a = torch.empty(b, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream(dev)
        _capi.check(lib.ga3c_select_actions(d_p.data_ptr(), d_u.data_ptr(), b, a, d_a.data_ptr(), st.cuda_stream),
                    "ga3c_select_actions")
        return d_a.cpu().numpy()
