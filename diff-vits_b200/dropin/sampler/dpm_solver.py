"""DPM-Solver / DPM-Solver++ - drop-in for `sampler/dpm_solver.py` of adelacvg/diff-vits.

Public surface kept: `NoiseScheduleVP(schedule, betas, alphas_cumprod, ...)`, `model_wrapper(...)`,
`DPM_Solver(model_fn, noise_schedule, algorithm_type, correcting_x0_fn, correcting_xt_fn, ...)` with
`.sample(x, steps, t_start, t_end, order, skip_type, method, lower_order_final, denoise_to_zero, solver_type, ...)`,
`.add_noise`, `.inverse`, plus `interpolate_fn` / `expand_dims`.

Fast path (what `model3.NaturalSpeech2.sample` uses, model3.py:1143-1158): algorithm 'dpmsolver++', method
'multistep', order <= 2, x_start model without guidance whose `model` is a B200-native denoiser -> the whole loop
(denoiser + updates) is one CUDA graph inside libdiffvits_b200 (`dv_sample`). Every other multistep setting runs the
generic per-step loop below over the same `model_fn` (orders 1-3, both algorithm types, 'dpmsolver'/'taylor').
The singlestep and adaptive solvers of the reference (dpm_solver.py:594-794, 956-1010) are not used by Diff-VITS and
raise NotImplementedError.
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

    # ------------------------------------------------------------------ sample
    def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
               lower_order_final=True, denoise_to_zero=False, solver_type="dpmsolver", atol=0.0078, rtol=0.05,
               return_intermediate=False):
        ns = self.noise_schedule
        t_0 = 1.0 / ns.total_N if t_end is None else t_end
        t_T = ns.T if t_start is None else t_start
        assert t_0 > 0 and t_T > 0, ("Time range needs to be greater than 0. For discrete-time DPMs, it needs to be "
                                     "in [1 / N, 1], where N is the length of betas array")
        if method != "multistep":
            if method in ("singlestep", "singlestep_fixed", "adaptive"):
                raise NotImplementedError(f"method={method!r}: only the multistep solver (the one Diff-VITS uses, "
                                          "model3.py:1157) is implemented in diffvits_b200")
            raise ValueError("Got wrong method {}".format(method))
        assert steps >= order
        self.last_run_native = False
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
