"""DPM-Solver / DPM-Solver++ - drop-in for `sampler/dpm_solver.py` of adelacvg/diff-vits.

Public surface kept: `NoiseScheduleVP(schedule, betas, alphas_cumprod, ...)`, `model_wrapper(...)`,
`DPM_Solver(model_fn, noise_schedule, algorithm_type, correcting_x0_fn, correcting_xt_fn, ...)` with
`.sample(x, steps, t_start, t_end, order, skip_type, method, lower_order_final, denoise_to_zero, solver_type, ...)`,
`.add_noise`, `.inverse`, plus `interpolate_fn` / `expand_dims`.

Fast path (what `model3.NaturalSpeech2.sample` uses, model3.py:1143-1158): algorithm 'dpmsolver++', method
'multistep', order <= 2, x_start model without guidance whose `model` is a B200-native denoiser -> the whole loop
(denoiser + updates) is one CUDA graph inside libdiffvits_b200 (`dv_sample`). Every other setting runs the generic
per-step Python loops below over the same `model_fn` (so still over the native denoiser call, one launch sequence per
model evaluation): multistep orders 1-3, the singlestep solvers ("DPM-Solver-fast", 'singlestep_fixed') and the
adaptive step-size solver (dpm_solver.py:594-794, 956-1010, 1214-1232), both algorithm types, 'dpmsolver'/'taylor'.

The singlestep formulas are written once for both parameterisations: with sgn = -1 (data prediction, "++") or +1
(noise prediction), e(h) = expm1(sgn h), D(s,t) = sigma_t/sigma_s or alpha_t/alpha_s and A(t) = alpha_t or sigma_t, the
first-order step is x_t = D x - A e(h) m_s and the higher orders add finite-difference corrections of the model values
at the intermediate times lambda_s + r_i h (Lu et al., DPM-Solver, Algorithms 1-2 / DPM-Solver++ eq. 9-11).
"""
from __future__ import annotations

import torch

from diffvits_b200.schedule import NoiseScheduleVP as _NoiseScheduleVP
from diffvits_b200.schedule import (dpmpp_multistep_table, expand_dims, get_time_steps, interpolate_fn,  # noqa: F401
                                    model_wrapper as _model_wrapper)

__all__ = ["NoiseScheduleVP", "model_wrapper", "DPM_Solver", "interpolate_fn", "expand_dims"]


class NoiseScheduleVP(_NoiseScheduleVP):
    """dpm_solver.py:6-167 flavour: 'discrete' | 'linear', with log-SNR clipping at -5.1 for discrete schedules."""

    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, continuous_beta_0=0.1,
                 continuous_beta_1=20.0, dtype=torch.float32):
        super().__init__(schedule, betas, alphas_cumprod, continuous_beta_0, continuous_beta_1, dtype,
                         clip_lambda=-5.1, allow_cosine=False)


def model_wrapper(model, noise_schedule, model_type="noise", model_kwargs={}, guidance_type="uncond", condition=None,
                  unconditional_condition=None, guidance_scale=1.0, classifier_fn=None, classifier_kwargs={}):
    return _model_wrapper(model, noise_schedule, model_type, model_kwargs, guidance_type, condition,
                          unconditional_condition, guidance_scale, classifier_fn, classifier_kwargs,
                          batched_coeffs=True)


class DPM_Solver:
    def __init__(self, model_fn, noise_schedule, algorithm_type="dpmsolver++", correcting_x0_fn=None,
                 correcting_xt_fn=None, thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995):
        assert algorithm_type in ["dpmsolver", "dpmsolver++"]
        self._wrapped = model_fn
        self.model = lambda x, t: model_fn(x, t.expand((x.shape[0])))
        self.noise_schedule = noise_schedule
        self.algorithm_type = algorithm_type
        self.correcting_x0_fn = (self.dynamic_thresholding_fn if correcting_x0_fn == "dynamic_thresholding"
                                 else correcting_x0_fn)
        self.correcting_xt_fn = correcting_xt_fn
        self.dynamic_thresholding_ratio = dynamic_thresholding_ratio
        self.thresholding_max_val = thresholding_max_val
        self.last_run_native = False

    # ------------------------------------------------------------------ model views
    def dynamic_thresholding_fn(self, x0, t):
        s = torch.quantile(torch.abs(x0).reshape((x0.shape[0], -1)), self.dynamic_thresholding_ratio, dim=1)
        s = expand_dims(torch.maximum(s, self.thresholding_max_val * torch.ones_like(s)), x0.dim())
        return torch.clamp(x0, -s, s) / s

    def noise_prediction_fn(self, x, t):
        return self.model(x, t)

    def data_prediction_fn(self, x, t):
        noise = self.noise_prediction_fn(x, t)
        alpha_t, sigma_t = self.noise_schedule.marginal_alpha(t), self.noise_schedule.marginal_std(t)
        x0 = (x - sigma_t * noise) / alpha_t
        if self.correcting_x0_fn is not None:
            x0 = self.correcting_x0_fn(x0, t)
        return x0

    def model_fn(self, x, t):
        return self.data_prediction_fn(x, t) if self.algorithm_type == "dpmsolver++" else self.noise_prediction_fn(x, t)

    def get_time_steps(self, skip_type, t_T, t_0, N, device):
        return get_time_steps(self.noise_schedule, skip_type, t_T, t_0, N, device)

    def denoise_to_zero_fn(self, x, s):
        return self.data_prediction_fn(x, s)

    # ------------------------------------------------------------------ multistep updates (orders 1..3)
    def dpm_solver_first_update(self, x, s, t, model_s=None, return_intermediate=False):
        ns = self.noise_schedule
        lam_s, lam_t = ns.marginal_lambda(s), ns.marginal_lambda(t)
        h = lam_t - lam_s
        if model_s is None:
            model_s = self.model_fn(x, s)
        if self.algorithm_type == "dpmsolver++":
            x_t = ns.marginal_std(t) / ns.marginal_std(s) * x - ns.marginal_alpha(t) * torch.expm1(-h) * model_s
        else:
            la_s, la_t = ns.marginal_log_mean_coeff(s), ns.marginal_log_mean_coeff(t)
            x_t = torch.exp(la_t - la_s) * x - ns.marginal_std(t) * torch.expm1(h) * model_s
        return (x_t, {"model_s": model_s}) if return_intermediate else x_t

    def multistep_dpm_solver_second_update(self, x, model_prev_list, t_prev_list, t, solver_type="dpmsolver"):
        if solver_type not in ["dpmsolver", "taylor"]:
            raise ValueError("'solver_type' must be either 'dpmsolver' or 'taylor', got {}".format(solver_type))
        ns = self.noise_schedule
        m1, m0 = model_prev_list[-2], model_prev_list[-1]
        t1, t0 = t_prev_list[-2], t_prev_list[-1]
        lam1, lam0, lam_t = ns.marginal_lambda(t1), ns.marginal_lambda(t0), ns.marginal_lambda(t)
        h0, h = lam0 - lam1, lam_t - lam0
        D1 = (1.0 / (h0 / h)) * (m0 - m1)
        if self.algorithm_type == "dpmsolver++":
            phi = torch.expm1(-h)
            base = ns.marginal_std(t) / ns.marginal_std(t0) * x - ns.marginal_alpha(t) * phi * m0
            if solver_type == "dpmsolver":
                return base - 0.5 * (ns.marginal_alpha(t) * phi) * D1
            return base + ns.marginal_alpha(t) * (phi / h + 1.0) * D1
        phi = torch.expm1(h)
        la0, la_t = ns.marginal_log_mean_coeff(t0), ns.marginal_log_mean_coeff(t)
        base = torch.exp(la_t - la0) * x - ns.marginal_std(t) * phi * m0
        if solver_type == "dpmsolver":
            return base - 0.5 * (ns.marginal_std(t) * phi) * D1
        return base - ns.marginal_std(t) * (phi / h - 1.0) * D1

    def multistep_dpm_solver_third_update(self, x, model_prev_list, t_prev_list, t, solver_type="dpmsolver"):
        ns = self.noise_schedule
        m2, m1, m0 = model_prev_list
        t2, t1, t0 = t_prev_list
        lam2, lam1, lam0, lam_t = (ns.marginal_lambda(t2), ns.marginal_lambda(t1), ns.marginal_lambda(t0),
                                   ns.marginal_lambda(t))
        h1, h0, h = lam1 - lam2, lam0 - lam1, lam_t - lam0
        r0, r1 = h0 / h, h1 / h
        D1_0 = (1.0 / r0) * (m0 - m1)
        D1_1 = (1.0 / r1) * (m1 - m2)
        D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
        D2 = (1.0 / (r0 + r1)) * (D1_0 - D1_1)
        if self.algorithm_type == "dpmsolver++":
            phi_1 = torch.expm1(-h)
            phi_2 = phi_1 / h + 1.0
            phi_3 = phi_2 / h - 0.5
            return (ns.marginal_std(t) / ns.marginal_std(t0) * x - ns.marginal_alpha(t) * phi_1 * m0
                    + ns.marginal_alpha(t) * phi_2 * D1 - ns.marginal_alpha(t) * phi_3 * D2)
        phi_1 = torch.expm1(h)
        phi_2 = phi_1 / h - 1.0
        phi_3 = phi_2 / h - 0.5
        la0, la_t = ns.marginal_log_mean_coeff(t0), ns.marginal_log_mean_coeff(t)
        return (torch.exp(la_t - la0) * x - ns.marginal_std(t) * phi_1 * m0 - ns.marginal_std(t) * phi_2 * D1
                - ns.marginal_std(t) * phi_3 * D2)

    # ------------------------------------------------------------------ singlestep updates (orders 2, 3) and adaptive solver
    def _param(self):
        """(sgn, D, A) of the exponential-integrator form shared by both parameterisations (module docstring)."""
        ns = self.noise_schedule
        if self.algorithm_type == "dpmsolver++":
            return (-1.0, lambda s, t: ns.marginal_std(t) / ns.marginal_std(s), ns.marginal_alpha)
        return (1.0, lambda s, t: torch.exp(ns.marginal_log_mean_coeff(t) - ns.marginal_log_mean_coeff(s)),
                ns.marginal_std)

    def get_orders_and_timesteps_for_singlestep_solver(self, steps, order, skip_type, t_T, t_0, device):
        """Split `steps` model evaluations into singlestep solvers of order <= `order` ("DPM-Solver-fast",
        dpm_solver.py:482-539): as many full-order steps as fit, the remainder as one or two lower-order steps."""
        if order not in (1, 2, 3):
            raise ValueError("'order' must be '1' or '2' or '3'.")
        if order == 1:
            orders, K = [1] * steps, 1
        elif order == 2:
            orders = [2] * (steps // 2) + ([1] if steps % 2 else [])
            K = len(orders)
        else:
            K = steps // 3 + 1
            tail = {0: [2, 1], 1: [1], 2: [2]}[steps % 3]
            orders = [3] * (K - len(tail)) + tail
        if skip_type == "logSNR":     # K outer intervals uniform in log-SNR, as in the DPM-Solver paper
            outer = self.get_time_steps(skip_type, t_T, t_0, K, device)
        else:                         # the outer grid is the subset of the `steps`-point grid where a solver step ends
            idx = torch.cumsum(torch.tensor([0] + orders), 0).to(device)
            outer = self.get_time_steps(skip_type, t_T, t_0, steps, device)[idx]
        return outer, orders

    def singlestep_dpm_solver_second_update(self, x, s, t, r1=0.5, model_s=None, return_intermediate=False,
                                            solver_type="dpmsolver"):
        if solver_type not in ["dpmsolver", "taylor"]:
            raise ValueError("'solver_type' must be either 'dpmsolver' or 'taylor', got {}".format(solver_type))
        r1 = 0.5 if r1 is None else r1
        ns = self.noise_schedule
        sgn, D, A = self._param()
        lam_s = ns.marginal_lambda(s)
        h = ns.marginal_lambda(t) - lam_s
        s1 = ns.inverse_lambda(lam_s + r1 * h)
        e1, e = torch.expm1(sgn * r1 * h), torch.expm1(sgn * h)
        if model_s is None:
            model_s = self.model_fn(x, s)
        model_s1 = self.model_fn(D(s, s1) * x - (A(s1) * e1) * model_s, s1)
        a_t = A(t)
        x_t = D(s, t) * x - (a_t * e) * model_s
        if solver_type == "dpmsolver":
            x_t = x_t - (0.5 / r1) * (a_t * e) * (model_s1 - model_s)
        else:
            x_t = x_t - sgn * (1.0 / r1) * (a_t * (e / h - sgn)) * (model_s1 - model_s)
        return (x_t, {"model_s": model_s, "model_s1": model_s1}) if return_intermediate else x_t

    def singlestep_dpm_solver_third_update(self, x, s, t, r1=1.0 / 3.0, r2=2.0 / 3.0, model_s=None, model_s1=None,
                                           return_intermediate=False, solver_type="dpmsolver"):
        if solver_type not in ["dpmsolver", "taylor"]:
            raise ValueError("'solver_type' must be either 'dpmsolver' or 'taylor', got {}".format(solver_type))
        r1 = 1.0 / 3.0 if r1 is None else r1
        r2 = 2.0 / 3.0 if r2 is None else r2
        ns = self.noise_schedule
        sgn, D, A = self._param()
        lam_s = ns.marginal_lambda(s)
        h = ns.marginal_lambda(t) - lam_s
        s1, s2 = ns.inverse_lambda(lam_s + r1 * h), ns.inverse_lambda(lam_s + r2 * h)
        e1, e2, e = torch.expm1(sgn * r1 * h), torch.expm1(sgn * r2 * h), torch.expm1(sgn * h)
        phi_22 = e2 / (r2 * h) - sgn
        phi_2 = e / h - sgn
        phi_3 = phi_2 / h - 0.5
        if model_s is None:
            model_s = self.model_fn(x, s)
        if model_s1 is None:
            model_s1 = self.model_fn(D(s, s1) * x - (A(s1) * e1) * model_s, s1)
        a_s2 = A(s2)
        x_s2 = D(s, s2) * x - (a_s2 * e2) * model_s - sgn * (r2 / r1) * (a_s2 * phi_22) * (model_s1 - model_s)
        model_s2 = self.model_fn(x_s2, s2)
        a_t = A(t)
        x_t = D(s, t) * x - (a_t * e) * model_s
        if solver_type == "dpmsolver":
            x_t = x_t - sgn * (1.0 / r2) * (a_t * phi_2) * (model_s2 - model_s)
        else:
            d10 = (1.0 / r1) * (model_s1 - model_s)
            d11 = (1.0 / r2) * (model_s2 - model_s)
            D1 = (r2 * d10 - r1 * d11) / (r2 - r1)
            D2 = 2.0 * (d11 - d10) / (r2 - r1)
            x_t = x_t - sgn * (a_t * phi_2) * D1 - (a_t * phi_3) * D2
        if return_intermediate:
            return x_t, {"model_s": model_s, "model_s1": model_s1, "model_s2": model_s2}
        return x_t

    def singlestep_dpm_solver_update(self, x, s, t, order, return_intermediate=False, solver_type="dpmsolver", r1=None,
                                     r2=None):
        if order == 1:
            return self.dpm_solver_first_update(x, s, t, return_intermediate=return_intermediate)
        if order == 2:
            return self.singlestep_dpm_solver_second_update(x, s, t, return_intermediate=return_intermediate,
                                                            solver_type=solver_type, r1=r1)
        if order == 3:
            return self.singlestep_dpm_solver_third_update(x, s, t, return_intermediate=return_intermediate,
                                                           solver_type=solver_type, r1=r1, r2=r2)
        raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))

    def dpm_solver_adaptive(self, x, order, t_T, t_0, h_init=0.05, atol=0.0078, rtol=0.05, theta=0.9, t_err=1e-5,
                            solver_type="dpmsolver"):
        """Adaptive step size in log-SNR (dpm_solver.py:956-1010): an embedded pair of orders (order-1, order) sharing
        their model evaluations estimates the local error E; a step is accepted when E <= 1 and the next step size is
        theta * h * E^(-1/order), never stepping past lambda(t_0)."""
        ns = self.noise_schedule
        if order == 2:
            def lower(x, s, t):
                return self.dpm_solver_first_update(x, s, t, return_intermediate=True)

            def higher(x, s, t, **kw):
                return self.singlestep_dpm_solver_second_update(x, s, t, r1=0.5, solver_type=solver_type, **kw)
        elif order == 3:
            def lower(x, s, t):
                return self.singlestep_dpm_solver_second_update(x, s, t, r1=1.0 / 3.0, return_intermediate=True,
                                                                solver_type=solver_type)

            def higher(x, s, t, **kw):
                return self.singlestep_dpm_solver_third_update(x, s, t, r1=1.0 / 3.0, r2=2.0 / 3.0,
                                                               solver_type=solver_type, **kw)
        else:
            raise ValueError("For adaptive step size solver, order must be 2 or 3, got {}".format(order))
        s = t_T * torch.ones((1,)).to(x)
        lam_s = ns.marginal_lambda(s)
        lam_end = ns.marginal_lambda(t_0 * torch.ones_like(s).to(x))
        h = h_init * torch.ones_like(s).to(x)
        x_prev = x
        self.adaptive_nfe = 0
        while torch.abs(s - t_0).mean() > t_err:
            t = ns.inverse_lambda(lam_s + h)
            x_lo, shared = lower(x, s, t)
            x_hi = higher(x, s, t, **shared)
            tol = torch.max(torch.ones_like(x) * atol, rtol * torch.max(torch.abs(x_lo), torch.abs(x_prev)))
            rms = torch.sqrt(torch.square(((x_hi - x_lo) / tol).reshape((x.shape[0], -1))).mean(dim=-1, keepdim=True))
            E = rms.max()
            if torch.all(E <= 1.0):
                x, s, x_prev = x_hi, t, x_lo
                lam_s = ns.marginal_lambda(s)
            h = torch.min(theta * h * torch.float_power(E, -1.0 / order).float(), lam_end - lam_s)
            self.adaptive_nfe += order
        return x

    def multistep_dpm_solver_update(self, x, model_prev_list, t_prev_list, t, order, solver_type="dpmsolver"):
        if order == 1:
            return self.dpm_solver_first_update(x, t_prev_list[-1], t, model_s=model_prev_list[-1])
        if order == 2:
            return self.multistep_dpm_solver_second_update(x, model_prev_list, t_prev_list, t, solver_type=solver_type)
        if order == 3:
            return self.multistep_dpm_solver_third_update(x, model_prev_list, t_prev_list, t, solver_type=solver_type)
        raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))

    # ------------------------------------------------------------------ noise utilities
    def add_noise(self, x, t, noise=None):
        alpha_t, sigma_t = self.noise_schedule.marginal_alpha(t), self.noise_schedule.marginal_std(t)
        if noise is None:
            noise = torch.randn((t.shape[0], *x.shape), device=x.device)
        x = x.reshape((-1, *x.shape))
        xt = expand_dims(alpha_t, x.dim()) * x + expand_dims(sigma_t, x.dim()) * noise
        return xt.squeeze(0) if t.shape[0] == 1 else xt

    def inverse(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
                lower_order_final=True, denoise_to_zero=False, solver_type="dpmsolver", atol=0.0078, rtol=0.05,
                return_intermediate=False):
        t_0 = 1.0 / self.noise_schedule.total_N if t_start is None else t_start
        t_T = self.noise_schedule.T if t_end is None else t_end
        assert t_0 > 0 and t_T > 0
        return self.sample(x, steps=steps, t_start=t_0, t_end=t_T, order=order, skip_type=skip_type, method=method,
                           lower_order_final=lower_order_final, denoise_to_zero=denoise_to_zero,
                           solver_type=solver_type, atol=atol, rtol=rtol, return_intermediate=return_intermediate)

    # ------------------------------------------------------------------ native fast path
    def _native(self, method, order, solver_type, denoise_to_zero, return_intermediate):
        dv = getattr(self._wrapped, "_dv", None)
        if (dv is None or self.algorithm_type != "dpmsolver++" or method != "multistep" or order > 2
                or solver_type not in ("dpmsolver", "taylor") or denoise_to_zero or return_intermediate
                or self.correcting_x0_fn is not None or self.correcting_xt_fn is not None
                or dv["model_type"] != "x_start" or dv["guidance_type"] != "uncond"
                or dv["noise_schedule"] is not self.noise_schedule):
            return None
        from diffvits_b200.native import find_native
        return find_native(dv["model"], dv["model_kwargs"])

    def _sample_singlestep(self, x, steps, t_T, t_0, order, skip_type, method, denoise_to_zero, solver_type, atol, rtol,
                           return_intermediate):
        """'adaptive', 'singlestep' and 'singlestep_fixed' branches of the reference's sample (dpm_solver.py:1166-1170,
        1214-1232): generic Python loops over `model_fn`."""
        if return_intermediate:
            assert method in ["multistep", "singlestep", "singlestep_fixed"], \
                "Cannot use adaptive solver when saving intermediate values"
        device = x.device
        intermediates = []
        step = 0
        with torch.no_grad():
            if method == "adaptive":
                x = self.dpm_solver_adaptive(x, order=order, t_T=t_T, t_0=t_0, atol=atol, rtol=rtol,
                                             solver_type=solver_type)
            else:
                if method == "singlestep":
                    outer, orders = self.get_orders_and_timesteps_for_singlestep_solver(
                        steps=steps, order=order, skip_type=skip_type, t_T=t_T, t_0=t_0, device=device)
                else:
                    orders = [order] * (steps // order)
                    outer = self.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=len(orders), device=device)
                for step, k in enumerate(orders):
                    s, t = outer[step], outer[step + 1]
                    # intermediate times of this step: k points of the same skip type inside [s, t]
                    lam = self.noise_schedule.marginal_lambda(
                        self.get_time_steps(skip_type=skip_type, t_T=s.item(), t_0=t.item(), N=k, device=device))
                    span = lam[-1] - lam[0]
                    r1 = (lam[1] - lam[0]) / span if k >= 2 else None
                    r2 = (lam[2] - lam[0]) / span if k >= 3 else None
                    x = self.singlestep_dpm_solver_update(x, s, t, k, solver_type=solver_type, r1=r1, r2=r2)
                    if self.correcting_xt_fn is not None:
                        x = self.correcting_xt_fn(x, t, step)
                    if return_intermediate:
                        intermediates.append(x)
            if denoise_to_zero:
                t = torch.ones((1,)).to(device) * t_0
                x = self.denoise_to_zero_fn(x, t)
                if self.correcting_xt_fn is not None:
                    x = self.correcting_xt_fn(x, t, step + 1)
                if return_intermediate:
                    intermediates.append(x)
        return (x, intermediates) if return_intermediate else x

    # ------------------------------------------------------------------ sample
    def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
               lower_order_final=True, denoise_to_zero=False, solver_type="dpmsolver", atol=0.0078, rtol=0.05,
               return_intermediate=False):
        ns = self.noise_schedule
        t_0 = 1.0 / ns.total_N if t_end is None else t_end
        t_T = ns.T if t_start is None else t_start
        assert t_0 > 0 and t_T > 0, ("Time range needs to be greater than 0. For discrete-time DPMs, it needs to be "
                                     "in [1 / N, 1], where N is the length of betas array")
        if method not in ("multistep", "singlestep", "singlestep_fixed", "adaptive"):
            raise ValueError("Got wrong method {}".format(method))
        self.last_run_native = False
        if method != "multistep":
            return self._sample_singlestep(x, steps, t_T, t_0, order, skip_type, method, denoise_to_zero, solver_type,
                                           atol, rtol, return_intermediate)
        assert steps >= order
        native = self._native(method, order, solver_type, denoise_to_zero, return_intermediate)
        if native is not None and x.is_cuda:
            from diffvits_b200.engine import SAMPLER_DPMPP_2M
            from diffvits_b200.native import tick_progress
            ts = self.get_time_steps(skip_type, t_T, t_0, steps, "cpu")
            t_in, coef = dpmpp_multistep_table(ns, ts, order, lower_order_final, solver_type)
            out = native.engine.sample(SAMPLER_DPMPP_2M, t_in, coef, x.to(torch.float32).contiguous().clone())
            tick_progress(native, steps)
            self.last_run_native = True
            return out.to(x.dtype)

        device = x.device
        intermediates = []
        with torch.no_grad():
            timesteps = self.get_time_steps(skip_type=skip_type, t_T=t_T, t_0=t_0, N=steps, device=device)
            assert timesteps.shape[0] - 1 == steps
            t = timesteps[0]
            t_hist, m_hist = [t], [self.model_fn(x, t)]
            if self.correcting_xt_fn is not None:
                x = self.correcting_xt_fn(x, t, 0)
            if return_intermediate:
                intermediates.append(x)
            for step in range(1, steps + 1):
                t = timesteps[step]
                if step < order:
                    step_order = step                      # warm-up with lower orders
                elif lower_order_final and steps < 10:
                    step_order = min(order, steps + 1 - step)
                else:
                    step_order = order
                x = self.multistep_dpm_solver_update(x, m_hist[-step_order:], t_hist[-step_order:], t, step_order,
                                                     solver_type=solver_type)
                if self.correcting_xt_fn is not None:
                    x = self.correcting_xt_fn(x, t, step)
                if return_intermediate:
                    intermediates.append(x)
                if step < steps:                           # the final model value is never needed
                    t_hist.append(t)
                    m_hist.append(self.model_fn(x, t))
                    del t_hist[:-order], m_hist[:-order]   # keep the newest `order` entries
            if denoise_to_zero:
                t = torch.ones((1,)).to(device) * t_0
                x = self.denoise_to_zero_fn(x, t)
                if self.correcting_xt_fn is not None:
                    x = self.correcting_xt_fn(x, t, steps + 1)
                if return_intermediate:
                    intermediates.append(x)
        return (x, intermediates) if return_intermediate else x
