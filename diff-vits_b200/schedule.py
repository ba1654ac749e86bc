"""VP noise schedule, model wrapper and the host-side coefficient tables of the multistep samplers.

Everything here is scalar / tiny-tensor fp32 work that runs ONCE per sampling call on the host; the per-element
updates run in csrc/elementwise.cu. The formulas restate the reference's samplers
(sampler/dpm_solver.py:6-167, 170-334, 547-580, 796-905; sampler/uni_pc.py:6-152, 155-233, 471-588) and are evaluated
with the same fp32 torch ops so the coefficients agree with the reference to the last bit or two.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

# --------------------------------------------------------------------------------------------------------------
# helpers


def expand_dims(v: torch.Tensor, dims: int) -> torch.Tensor:
    """[N] -> [N, 1, ..., 1] with `dims` dimensions (reference: dpm_solver.py:1295-1304)."""
    return v[(...,) + (None,) * (dims - 1)]


def interpolate_fn(x: torch.Tensor, xp: torch.Tensor, yp: torch.Tensor) -> torch.Tensor:
    """Piecewise-linear y(x) through the knots (xp, yp), linear extrapolation outside; differentiable in x.

    x: [N, C], xp/yp: [C, K] (or [1, K], broadcast). Same contract and the same evaluation order
    y0 + (x - x0) * (y1 - y0) / (x1 - x0) as the reference (dpm_solver.py:1253-1292); the segment is located with
    searchsorted instead of the reference's sort/argmin.
    """
    n, k = x.shape[0], xp.shape[1]
    xp_b = xp.expand(x.shape[1], k).contiguous()
    yp_b = yp.expand(x.shape[1], k).contiguous()
    xt = x.transpose(0, 1).contiguous()                      # [C, N]
    below = torch.searchsorted(xp_b, xt, right=False)        # number of knots strictly below x
    start = torch.where(below == 0, torch.zeros_like(below), torch.where(below == k, below - 2, below - 1))
    end = start + 1
    x0, x1 = torch.gather(xp_b, 1, start), torch.gather(xp_b, 1, end)
    y0, y1 = torch.gather(yp_b, 1, start), torch.gather(yp_b, 1, end)
    y = y0 + (xt - x0) * (y1 - y0) / (x1 - x0)
    return y.transpose(0, 1).reshape(n, x.shape[1])


# --------------------------------------------------------------------------------------------------------------
# noise schedule


class NoiseScheduleVP:
    """Forward VP-SDE schedule: log alpha_t, alpha_t, sigma_t, lambda_t = log(alpha_t/sigma_t) and its inverse.

    schedule='discrete' (betas or alphas_cumprod of a discrete-time DPM; log alpha_t is piecewise linear over
    t_n = n/N), 'linear' (continuous VPSDE) or 'cosine'. `clip_lambda` applies the DPM-Solver variant's
    log-SNR clipping near t=T (dpm_solver.py:114-125); the UniPC variant (uni_pc.py:6-62) does not clip.
    """

    def __init__(self, schedule: str = "discrete", betas=None, alphas_cumprod=None, continuous_beta_0: float = 0.1,
                 continuous_beta_1: float = 20.0, dtype=torch.float32, *, clip_lambda: Optional[float] = None,
                 allow_cosine: bool = True):
        allowed = ["discrete", "linear"] + (["cosine"] if allow_cosine else [])
        if schedule not in allowed:
            raise ValueError("Unsupported noise schedule {}. The schedule needs to be {}".format(
                schedule, " or ".join(repr(a) for a in allowed)))
        self.schedule = schedule
        self.T = 1.0
        if schedule == "discrete":
            if betas is not None:
                log_alphas = 0.5 * torch.log(1 - betas).cumsum(dim=0)
            else:
                assert alphas_cumprod is not None
                log_alphas = 0.5 * torch.log(alphas_cumprod)
            if clip_lambda is not None:
                log_alphas = self.numerical_clip_alpha(log_alphas, clip_lambda)
            self.log_alpha_array = log_alphas.reshape((1, -1)).to(dtype=dtype)
            self.total_N = self.log_alpha_array.shape[1]
            self.t_array = torch.linspace(0.0, 1.0, self.total_N + 1)[1:].reshape((1, -1)).to(dtype=dtype)
        else:
            self.total_N = 1000
            self.beta_0, self.beta_1 = continuous_beta_0, continuous_beta_1
            self.cosine_s = 0.008
            self.cosine_beta_max = 999.0
            self.cosine_t_max = (math.atan(self.cosine_beta_max * (1.0 + self.cosine_s) / math.pi) * 2.0
                                 * (1.0 + self.cosine_s) / math.pi - self.cosine_s)
            self.cosine_log_alpha_0 = math.log(math.cos(self.cosine_s / (1.0 + self.cosine_s) * math.pi / 2.0))
            if schedule == "cosine":
                self.T = 0.9946

    @staticmethod
    def numerical_clip_alpha(log_alphas: torch.Tensor, clipped_lambda: float = -5.1) -> torch.Tensor:
        """Drop the trailing entries whose half-logSNR is below `clipped_lambda`."""
        log_sigmas = 0.5 * torch.log(1.0 - torch.exp(2.0 * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = int(torch.searchsorted(torch.flip(lambs, [0]), clipped_lambda))
        return log_alphas[:-idx] if idx > 0 else log_alphas

    def marginal_log_mean_coeff(self, t: torch.Tensor) -> torch.Tensor:
        if self.schedule == "discrete":
            return interpolate_fn(t.reshape((-1, 1)), self.t_array.to(t.device),
                                  self.log_alpha_array.to(t.device)).reshape((-1))
        if self.schedule == "linear":
            return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        log_alpha = torch.log(torch.cos((t + self.cosine_s) / (1.0 + self.cosine_s) * math.pi / 2.0))
        return log_alpha - self.cosine_log_alpha_0

    def marginal_alpha(self, t):
        return torch.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t):
        return torch.sqrt(1.0 - torch.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t):
        log_mean = self.marginal_log_mean_coeff(t)
        log_std = 0.5 * torch.log(1.0 - torch.exp(2.0 * log_mean))
        return log_mean - log_std

    def inverse_lambda(self, lamb):
        if self.schedule == "linear":
            tmp = 2.0 * (self.beta_1 - self.beta_0) * torch.logaddexp(-2.0 * lamb, torch.zeros((1,)).to(lamb))
            delta = self.beta_0 ** 2 + tmp
            return tmp / (torch.sqrt(delta) + self.beta_0) / (self.beta_1 - self.beta_0)
        if self.schedule == "discrete":
            log_alpha = -0.5 * torch.logaddexp(torch.zeros((1,)).to(lamb.device), -2.0 * lamb)
            t = interpolate_fn(log_alpha.reshape((-1, 1)), torch.flip(self.log_alpha_array.to(lamb.device), [1]),
                               torch.flip(self.t_array.to(lamb.device), [1]))
            return t.reshape((-1,))
        log_alpha = -0.5 * torch.logaddexp(-2.0 * lamb, torch.zeros((1,)).to(lamb))
        return (torch.arccos(torch.exp(log_alpha + self.cosine_log_alpha_0)) * 2.0 * (1.0 + self.cosine_s) / math.pi
                - self.cosine_s)


# --------------------------------------------------------------------------------------------------------------
# model wrapper


def model_wrapper(model: Callable, noise_schedule: NoiseScheduleVP, model_type: str = "noise",
                  model_kwargs: Optional[dict] = None, guidance_type: str = "uncond", condition=None,
                  unconditional_condition=None, guidance_scale: float = 1.0, classifier_fn=None,
                  classifier_kwargs: Optional[dict] = None, *, batched_coeffs: bool = True) -> Callable:
    """Wrap `model(x, t_input, **model_kwargs)` into the continuous-time noise predictor the solvers consume.

    Same contract as the reference (dpm_solver.py:170-334 / uni_pc.py:155-233). The returned function carries a
    `_dv` attribute with the wrapping recipe so the solvers can recognise a B200-native denoiser and run the whole
    loop natively. `batched_coeffs=True` broadcasts alpha_t / sigma_t over the batch (the reference's UniPC wrapper
    omits this and only works for batch 1: uni_pc.py:189-191).
    """
    model_kwargs = {} if model_kwargs is None else model_kwargs
    classifier_kwargs = {} if classifier_kwargs is None else classifier_kwargs
    assert model_type in ["noise", "x_start", "v", "score"]
    assert guidance_type in ["uncond", "classifier", "classifier-free"]
    ns = noise_schedule

    def model_time(t_continuous):
        if ns.schedule == "discrete":
            return (t_continuous - 1.0 / ns.total_N) * ns.total_N
        return t_continuous

    def bc(v, x):
        return expand_dims(v, x.dim()) if batched_coeffs else v

    def noise_pred_fn(x, t_continuous, cond=None):
        t_input = model_time(t_continuous)
        out = model(x, t_input, **model_kwargs) if cond is None else model(x, t_input, cond, **model_kwargs)
        if model_type == "noise":
            return out
        if model_type == "x_start":
            alpha_t, sigma_t = ns.marginal_alpha(t_continuous), ns.marginal_std(t_continuous)
            return (x - bc(alpha_t, x) * out) / bc(sigma_t, x)
        if model_type == "v":
            alpha_t, sigma_t = ns.marginal_alpha(t_continuous), ns.marginal_std(t_continuous)
            return bc(alpha_t, x) * out + bc(sigma_t, x) * x
        sigma_t = ns.marginal_std(t_continuous)      # "score"
        return -bc(sigma_t, x) * out

    def cond_grad_fn(x, t_input):
        with torch.enable_grad():
            x_in = x.detach().requires_grad_(True)
            log_prob = classifier_fn(x_in, t_input, condition, **classifier_kwargs)
            return torch.autograd.grad(log_prob.sum(), x_in)[0]

    def model_fn(x, t_continuous):
        if guidance_type == "uncond":
            return noise_pred_fn(x, t_continuous)
        if guidance_type == "classifier":
            assert classifier_fn is not None
            grad = cond_grad_fn(x, model_time(t_continuous))
            sigma_t = ns.marginal_std(t_continuous)
            return noise_pred_fn(x, t_continuous) - guidance_scale * bc(sigma_t, x) * grad
        if guidance_scale == 1.0 or unconditional_condition is None:
            return noise_pred_fn(x, t_continuous, cond=condition)
        x_in, t_in = torch.cat([x] * 2), torch.cat([t_continuous] * 2)
        c_in = torch.cat([unconditional_condition, condition])
        noise_uncond, noise = noise_pred_fn(x_in, t_in, cond=c_in).chunk(2)
        return noise_uncond + guidance_scale * (noise - noise_uncond)

    model_fn._dv = dict(model=model, noise_schedule=ns, model_type=model_type, model_kwargs=model_kwargs,
                        guidance_type=guidance_type)
    return model_fn


# --------------------------------------------------------------------------------------------------------------
# time grids


def get_time_steps(ns: NoiseScheduleVP, skip_type: str, t_T: float, t_0: float, N: int, device) -> torch.Tensor:
    """N+1 time points from t_T down to t_0 (reference: dpm_solver.py:453-480)."""
    if skip_type == "logSNR":
        lam_T = ns.marginal_lambda(torch.tensor(t_T).to(device))
        lam_0 = ns.marginal_lambda(torch.tensor(t_0).to(device))
        return ns.inverse_lambda(torch.linspace(lam_T.cpu().item(), lam_0.cpu().item(), N + 1).to(device))
    if skip_type == "time_uniform":
        return torch.linspace(t_T, t_0, N + 1).to(device)
    if skip_type == "time_quadratic":
        return torch.linspace(t_T ** 0.5, t_0 ** 0.5, N + 1).pow(2).to(device)
    raise ValueError(
        "Unsupported skip_type {}, need to be 'logSNR' or 'time_uniform' or 'time_quadratic'".format(skip_type))


def model_input_times(ns: NoiseScheduleVP, timesteps: torch.Tensor) -> torch.Tensor:
    if ns.schedule == "discrete":
        return (timesteps - 1.0 / ns.total_N) * ns.total_N
    return timesteps


# --------------------------------------------------------------------------------------------------------------
# coefficient tables for the native loop (one row of 8 floats per solver step; layout in include/diffvits_b200.h)


_TABLE_CACHE: dict = {}


def _schedule_fingerprint(ns: "NoiseScheduleVP"):
    """Content key of a schedule: the reference code (model3.py:1143, 1177) builds a NEW NoiseScheduleVP per sample() call,
    so the memo below cannot key on object identity."""
    fp = ns.__dict__.get("_dv_fingerprint")
    if fp is None:
        if ns.schedule == "discrete":
            body = hash(ns.log_alpha_array.detach().to("cpu", torch.float32).contiguous().numpy().tobytes())
        else:
            body = (getattr(ns, "beta_0", None), getattr(ns, "beta_1", None))
        fp = ns.__dict__["_dv_fingerprint"] = (ns.schedule, float(ns.T), int(ns.total_N), body)
    return fp


def _table_memo(ns: "NoiseScheduleVP", kind: str, timesteps: torch.Tensor, args: tuple, build):
    """Coefficient tables depend on (schedule, time points, solver settings) only: ~3 ms of scalar torch ops on the host per
    build, i.e. 1.5 % of a C2 job and 5 % of a 10-step UniPC job when redone per `sample()` call. They are memoised by
    schedule content; the returned CPU tensors are shared and must be treated as read-only (the engine copies them)."""
    key = (_schedule_fingerprint(ns), kind, tuple(timesteps.detach().to("cpu", torch.float32).tolist()), args)
    hit = _TABLE_CACHE.get(key)
    if hit is None:
        if len(_TABLE_CACHE) >= 64:
            _TABLE_CACHE.clear()
        hit = _TABLE_CACHE[key] = build()
    return hit


def dpmpp_multistep_table(ns: NoiseScheduleVP, timesteps: torch.Tensor, order: int, lower_order_final: bool,
                          solver_type: str = "dpmsolver") -> Tuple[torch.Tensor, torch.Tensor]:
    return _table_memo(ns, "dpmpp", timesteps, (order, bool(lower_order_final), solver_type),
                       lambda: _dpmpp_multistep_table(ns, timesteps, order, lower_order_final, solver_type))


def _dpmpp_multistep_table(ns: NoiseScheduleVP, timesteps: torch.Tensor, order: int, lower_order_final: bool,
                           solver_type: str = "dpmsolver") -> Tuple[torch.Tensor, torch.Tensor]:
    """DPM-Solver++ (data prediction) multistep, order <= 2:  x_i = A x + B m_{-1} + C m_{-2}.

    order 1 (dpm_solver.py:573-576):  A = sigma_t/sigma_s, B = -alpha_t*expm1(-h)
    order 2 (dpm_solver.py:820-831):  D1 = (m_{-1} - m_{-2})/r0 ;
        'dpmsolver': x = A x - alpha_t*phi*m_{-1} - 0.5*alpha_t*phi*D1      'taylor': ... + alpha_t*(phi/h + 1)*D1
    Returns (t_in [steps], coef [steps, 8]) on the CPU in fp32.
    """
    assert order in (1, 2) and solver_type in ("dpmsolver", "taylor")
    ts = timesteps.detach().to("cpu", torch.float32)
    steps = ts.shape[0] - 1
    lam, log_alpha = ns.marginal_lambda(ts), ns.marginal_log_mean_coeff(ts)
    sigma, alpha = ns.marginal_std(ts), torch.exp(log_alpha)
    coef = torch.zeros(steps, 8, dtype=torch.float32)
    for i in range(1, steps + 1):
        if i < order:
            step_order = i
        elif lower_order_final and steps < 10:
            step_order = min(order, steps + 1 - i)
        else:
            step_order = order
        h = lam[i] - lam[i - 1]
        phi = torch.expm1(-h)
        a = sigma[i] / sigma[i - 1]
        if step_order == 1:
            b, c, has_prev = -(alpha[i] * phi), torch.zeros(()), 0.0
        else:
            r0 = (lam[i - 1] - lam[i - 2]) / h
            if solver_type == "dpmsolver":
                k = -0.5 * (alpha[i] * phi) / r0          # multiplies (m_{-1} - m_{-2})
            else:
                k = (alpha[i] * (phi / h + 1.0)) / r0
            b, c, has_prev = -(alpha[i] * phi) + k, -k, 1.0
        coef[i - 1, 0], coef[i - 1, 1], coef[i - 1, 2], coef[i - 1, 3] = a, b, c, has_prev
        coef[i - 1, 4], coef[i - 1, 5] = alpha[i - 1], sigma[i - 1]
    return model_input_times(ns, ts[:-1]).contiguous(), coef


def unipc_bh_rhos(hh: torch.Tensor, rks: Sequence[torch.Tensor], variant: str):
    """Solve the UniPC order conditions (uni_pc.py:500-545) for the predictor / corrector weights."""
    order = len(rks)
    h_phi_1 = torch.expm1(hh)
    h_phi_k = h_phi_1 / hh - 1
    if variant == "bh1":
        b_h = hh
    elif variant == "bh2":
        b_h = torch.expm1(hh)
    else:
        raise NotImplementedError()
    rk = torch.stack(list(rks))
    rows, rhs, fact = [], [], 1
    for i in range(1, order + 1):
        rows.append(torch.pow(rk, i - 1))
        rhs.append(h_phi_k * fact / b_h)
        fact *= i + 1
        h_phi_k = h_phi_k / hh - 1 / fact
    R, b = torch.stack(rows), torch.stack(rhs)
    if order == 1:
        rhos_p, rhos_c = None, torch.tensor([0.5], dtype=hh.dtype)
    else:
        rhos_p = torch.tensor([0.5], dtype=hh.dtype) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
        rhos_c = torch.linalg.solve(R, b)
    return h_phi_1, b_h, rhos_p, rhos_c


def unipc_multistep_table(ns: NoiseScheduleVP, timesteps: torch.Tensor, order: int, lower_order_final: bool,
                          variant: str = "bh2") -> Tuple[torch.Tensor, torch.Tensor]:
    return _table_memo(ns, "unipc", timesteps, (order, bool(lower_order_final), variant),
                       lambda: _unipc_multistep_table(ns, timesteps, order, lower_order_final, variant))


def _unipc_multistep_table(ns: NoiseScheduleVP, timesteps: torch.Tensor, order: int, lower_order_final: bool,
                           variant: str = "bh2") -> Tuple[torch.Tensor, torch.Tensor]:
    """UniPC (data prediction, B(h) variant) multistep, order <= 2 (uni_pc.py:471-588, 606-658):

        xbar   = a x + b m_{-1}                         a = sigma_t/sigma_s, b = -alpha_t*h_phi_1
        x_pred = xbar + cp (m_{-2} - m_{-1})            cp = -alpha_t*B_h*rho_p/r
        base   = xbar + cc (m_{-2} - m_{-1})            cc = -alpha_t*B_h*rho_c0/r
        x      = base + d (m_t - m_{-1})                d  = -alpha_t*B_h*rho_c_last         (corrector steps)
    """
    assert order in (1, 2)
    ts = timesteps.detach().to("cpu", torch.float32)
    steps = ts.shape[0] - 1
    lam, sigma = ns.marginal_lambda(ts), ns.marginal_std(ts)
    alpha = torch.exp(ns.marginal_log_mean_coeff(ts))
    coef = torch.zeros(steps, 8, dtype=torch.float32)
    for i in range(1, steps + 1):
        if i < order:
            step_order = i
        else:
            step_order = min(order, steps + 1 - i) if lower_order_final else order
        use_corr = i < steps
        h = lam[i] - lam[i - 1]
        rks = []
        if step_order == 2:
            rks.append((lam[i - 2] - lam[i - 1]) / h)
        rks.append(torch.tensor(1.0))
        h_phi_1, b_h, rhos_p, rhos_c = unipc_bh_rhos(-h, rks, variant)
        a = sigma[i] / sigma[i - 1]
        b = -(alpha[i] * h_phi_1)
        cp = cc = torch.zeros(())
        if step_order == 2:
            cp = -(alpha[i] * b_h) * rhos_p[0] / rks[0]
            cc = -(alpha[i] * b_h) * rhos_c[0] / rks[0]
        d = -(alpha[i] * b_h) * rhos_c[-1]
        row = coef[i - 1]
        row[0], row[1], row[2], row[3], row[4] = a, b, cp, cc, d
        row[5], row[6] = float(step_order == 2), float(use_corr)
    return model_input_times(ns, ts[:-1]).contiguous(), coef
