"""torch-CPU fp32 restatement of the same graph.  TEST INFRASTRUCTURE ONLY.

Two jobs:
  1. an independent second opinion on layout / SAME padding / autodiff for `oracle_np.py`
     (torch's conv2d and autograd share no code with the numpy im2col restatement);
  2. the multi-threaded CPU baseline that `bench.py --impl reference` and the `cpu_baseline`
     leg time: "CPU restatement of the reference TF graph (TensorFlow not installable)".

It is NOT a valid second opinion for RMSProp: torch.optim.RMSprop has eps outside the sqrt and a
zero-initialised square average; the optimizer here is hand-written with TF semantics.
Reference call sites: NetworkVP.py:212-228, NetworkDNav.py:81-90,256-269,
NetworkVP_discrate.py:60-85,99-105.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle_np as onp


class TorchNetworkVP:
    def __init__(self, params: dict, *, dtype=torch.float32, rho=0.99, mu=0.0, eps=0.1,
                 log_eps=1e-6, min_policy=0.0):
        self.dtype = dtype
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in params.items()}
        self.ms = {k: torch.ones_like(v) for k, v in self.p.items()}
        self.mom = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.rho, self.mu, self.eps, self.log_eps, self.min_policy = rho, mu, eps, log_eps, min_policy

    # -- forward ---------------------------------------------------------------------
    def _trunk(self, x):
        b = x.shape[0]
        x = x.reshape(b, onp.H, onp.W, onp.C).permute(0, 3, 1, 2)                  # NHWC -> NCHW
        w11 = self.p["conv11/w:0"].permute(3, 2, 0, 1)                             # HWIO -> OIHW
        w12 = self.p["conv12/w:0"].permute(3, 2, 0, 1)
        x = F.pad(x, (onp.P1_LO, onp.P1_HI, onp.P1_LO, onp.P1_HI))
        n1 = F.relu(F.conv2d(x, w11, self.p["conv11/b:0"], stride=onp.C1_S))
        n1 = F.pad(n1, (onp.P2_LO, onp.P2_HI, onp.P2_LO, onp.P2_HI))               # asymmetric (1, 2)
        n2 = F.relu(F.conv2d(n1, w12, self.p["conv12/b:0"], stride=onp.C2_S))
        flat = n2.permute(0, 2, 3, 1).reshape(b, onp.FLAT)                         # (h, w, c) flatten
        return F.relu(flat @ self.p["dense1/w:0"] + self.p["dense1/b:0"])

    def heads(self, x):
        d1 = self._trunk(x)
        v = (d1 @ self.p["logits_v/w:0"] + self.p["logits_v/b:0"])[:, 0]
        z = d1 @ self.p["logits_p/w:0"] + self.p["logits_p/b:0"]
        s = torch.softmax(z, dim=1)
        p = (s + self.min_policy) / (1.0 + self.min_policy * z.shape[1])
        return p, v

    @torch.no_grad()
    def predict_p_and_v(self, x):
        p, v = self.heads(torch.as_tensor(x, dtype=self.dtype))
        return p.numpy(), v.numpy()

    # -- loss / train ------------------------------------------------------------------
    def loss(self, x, y_r, a, beta):
        p, v = self.heads(torch.as_tensor(x, dtype=self.dtype))
        y_r = torch.as_tensor(y_r, dtype=self.dtype)
        a = torch.as_tensor(a, dtype=self.dtype)
        eps = torch.tensor(self.log_eps, dtype=self.dtype)
        sel = (p * a).sum(dim=1)
        cost_p_1 = torch.log(torch.maximum(sel, eps)) * (y_r - v.detach())
        cost_p_2 = -beta * (torch.log(torch.maximum(p, eps)) * p).sum(dim=1)
        c1, c2 = cost_p_1.sum(), cost_p_2.sum()
        cost_p = -(c1 + c2)
        cost_v = 0.5 * ((y_r - v) ** 2).sum()
        return dict(cost_p_1=c1, cost_p_2=c2, cost_p=cost_p, cost_v=cost_v, cost_all=cost_p + cost_v)

    def grads(self, x, y_r, a, beta):
        for t in self.p.values():
            t.grad = None
        losses = self.loss(x, y_r, a, beta)
        losses["cost_all"].backward()
        return ({k: float(v.detach()) for k, v in losses.items()},
                {k: t.grad.detach().numpy().copy() for k, t in self.p.items()})

    def train(self, x, y_r, a, lr, beta):
        """One A5 step: forward, autograd backward, TF-semantics RMSProp (eps inside sqrt, ms0 = 1)."""
        for t in self.p.values():
            t.grad = None
        self.loss(x, y_r, a, beta)["cost_all"].backward()
        with torch.no_grad():
            for k, w in self.p.items():
                g = w.grad
                self.ms[k].mul_(self.rho).addcmul_(g, g, value=1.0 - self.rho)
                self.mom[k].mul_(self.mu).add_(lr * g / torch.sqrt(self.ms[k] + self.eps))
                w.sub_(self.mom[k])
