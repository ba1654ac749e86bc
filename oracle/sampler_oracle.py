"""ORACLE (test infrastructure, not product code): plain-torch restatement of the two sampler loops Diff-VITS runs,
independent of diff-vits_b200/schedule.py. Only tests/, smoke() and bench.py's CPU-baseline legs may import it.

 * DPM-Solver++ 2M multistep  (sampler/dpm_solver.py:1171-1213 with :547-580 and :796-852)
 * UniPC-B(h2) multistep, predict_x0 (sampler/uni_pc.py:606-658 with :471-588)
both over the 'discrete' VP schedule (dpm_solver.py:88-167 / uni_pc.py:38-134) and the x_start model wrapper
(dpm_solver.py:268-292). Pinned against the live reference in tests/test_oracle.py and against tests/golden/.

`model(x, t_in)` is the x0-predicting denoiser called with the model time t_in = (t - 1/N) * N (dpm_solver.py:278).
`roundtrip=True` reproduces the reference's x0 -> eps -> x0 conversion (dpm_solver.py:290-292 then :437-439).
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch

Tensor = torch.Tensor


class DiscreteVP:
    """log alpha_t piecewise linear over t_n = n/N, n = 1..N (dpm_solver.py:100-112, 127-154)."""

    def __init__(self, betas: Tensor, clip_lambda: Optional[float] = None):
        la = 0.5 * torch.log(1 - betas.to(torch.float32)).cumsum(dim=0)
        if clip_lambda is not None:                       # dpm_solver.py:114-125
            lam = la - 0.5 * torch.log(1.0 - torch.exp(2.0 * la))
            drop = int((lam < clip_lambda).sum())
            if drop > 0:
                la = la[:-drop]
        self.la = la.to(torch.float32)
        self.N = self.la.numel()
        self.t = torch.linspace(0.0, 1.0, self.N + 1)[1:].to(torch.float32)

    def log_alpha(self, t: Tensor) -> Tensor:
        t = t.to(torch.float32)
        idx = int(torch.searchsorted(self.t, t.reshape(1), right=False)[0])
        lo = 0 if idx == 0 else (self.N - 2 if idx == self.N else idx - 1)
        x0, x1, y0, y1 = self.t[lo], self.t[lo + 1], self.la[lo], self.la[lo + 1]
        return y0 + (t - x0) * (y1 - y0) / (x1 - x0)

    def alpha(self, t):
        return torch.exp(self.log_alpha(t))

    def sigma(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.log_alpha(t)))

    def lam(self, t):
        la = self.log_alpha(t)
        return la - 0.5 * torch.log(1.0 - torch.exp(2.0 * la))

    def t_in(self, t):
        return (t - 1.0 / self.N) * self.N


def _x0(ns: DiscreteVP, model: Callable, x: Tensor, t: Tensor, roundtrip: bool) -> Tensor:
    out = model(x, ns.t_in(t).reshape(1).expand(x.shape[0]))
    if not roundtrip:
        return out
    a, s = ns.alpha(t), ns.sigma(t)
    eps = (x - a * out) / s
    return (x - s * eps) / a


def dpmpp_2m(model: Callable, betas: Tensor, x: Tensor, steps: int, order: int = 2, roundtrip: bool = True,
             outs: Optional[List[Tensor]] = None) -> Tensor:
    ns = DiscreteVP(betas, clip_lambda=-5.1)
    ts = torch.linspace(1.0, 1.0 / ns.N, steps + 1)                  # time_uniform, dpm_solver.py:473-474
    hist_m = [_x0(ns, model, x, ts[0], roundtrip)]
    if outs is not None:
        outs.append(hist_m[-1])
    for i in range(1, steps + 1):
        so = i if i < order else (min(order, steps + 1 - i) if steps < 10 else order)   # :1198-1201
        s, t = ts[i - 1], ts[i]
        h = ns.lam(t) - ns.lam(s)
        phi = torch.expm1(-h)
        x_new = ns.sigma(t) / ns.sigma(s) * x - ns.alpha(t) * phi * hist_m[-1]           # :573-576
        if so == 2:
            h0 = ns.lam(s) - ns.lam(ts[i - 2])
            d1 = (1.0 / (h0 / h)) * (hist_m[-1] - hist_m[-2])                            # :813-817
            x_new = x_new - 0.5 * (ns.alpha(t) * phi) * d1                               # :820-825
        x = x_new
        if i < steps:
            hist_m.append(_x0(ns, model, x, t, roundtrip))
            hist_m = hist_m[-order:]
            if outs is not None:
                outs.append(hist_m[-1])
    return x


def unipc_bh2(model: Callable, betas: Tensor, x: Tensor, steps: int, order: int = 2, roundtrip: bool = True,
              outs: Optional[List[Tensor]] = None) -> Tensor:
    ns = DiscreteVP(betas, clip_lambda=None)
    ts = torch.linspace(1.0, 1.0 / ns.N, steps + 1)
    hist_m = [_x0(ns, model, x, ts[0], roundtrip)]
    if outs is not None:
        outs.append(hist_m[-1])
    for i in range(1, steps + 1):
        so = i if i < order else min(order, steps + 1 - i)                               # uni_pc.py:636-639
        use_corr = i < steps                                                             # :640-644
        s, t = ts[i - 1], ts[i]
        h = ns.lam(t) - ns.lam(s)
        hh = -h
        h_phi_1 = torch.expm1(hh)                                                        # :503
        b_h = torch.expm1(hh)                                                            # bh2 :511-512
        m0 = hist_m[-1]
        xbar = ns.sigma(t) / ns.sigma(s) * x - ns.alpha(t) * h_phi_1 * m0                # :548-551
        h_phi_k = h_phi_1 / hh - 1
        if so == 2:
            rk = (ns.lam(ts[i - 2]) - ns.lam(s)) / h                                     # :488-493
            d1 = (hist_m[-2] - m0) / rk
            b1 = h_phi_k * 1 / b_h                                                       # :516-520
            h_phi_k2 = h_phi_k / hh - 1 / 2
            b2 = h_phi_k2 * 2 / b_h
            R = torch.stack([torch.stack([torch.ones(()), torch.ones(())]), torch.stack([rk, torch.ones(())])])
            rhos_c = torch.linalg.solve(R, torch.stack([b1, b2]))                        # :544
            x_pred = xbar - ns.alpha(t) * b_h * (0.5 * d1)                               # rhos_p = [0.5], :531-532
        else:
            d1 = None
            rhos_c = torch.tensor([0.5])                                                 # :541-542
            x_pred = xbar
        if use_corr:
            m_t = _x0(ns, model, x_pred, t, roundtrip)
            corr = rhos_c[0] * d1 if d1 is not None else 0
            x = xbar - ns.alpha(t) * b_h * (corr + rhos_c[-1] * (m_t - m0))              # :563-567
            hist_m.append(m_t)
            hist_m = hist_m[-order:]
            if outs is not None:
                outs.append(m_t)
        else:
            x = x_pred
    return x
