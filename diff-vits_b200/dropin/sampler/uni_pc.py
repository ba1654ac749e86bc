"""UniPC - drop-in for `sampler/uni_pc.py` of adelacvg/diff-vits.

Public surface kept: `NoiseScheduleVP` ('discrete' | 'linear' | 'cosine', no log-SNR clipping), `model_wrapper`,
`UniPC(model_fn, noise_schedule, algorithm_type, correcting_x0_fn, correcting_xt_fn, ..., variant)` with
`.sample(x, steps, t_start, t_end, order, skip_type, method, lower_order_final, denoise_to_zero, ...)`.

Fast path (model3.py:1177-1192: variant 'bh2', order 2, multistep, x_start model): the predictor / corrector
updates and the denoiser run as one CUDA graph inside libdiffvits_b200. Other multistep settings (bh1, order 3,
noise prediction, correctors, `vary_coeff`) run the generic per-step loop below over the same denoiser.

Difference from the reference on purpose: `model_wrapper` broadcasts alpha_t / sigma_t over the batch, so batch > 1
works (the reference raises a shape error there, uni_pc.py:189-191); for batch 1 the results are identical.
"""
from __future__ import annotations

import torch

from diffvits_b200.schedule import NoiseScheduleVP as _NoiseScheduleVP
from diffvits_b200.schedule import (expand_dims, get_time_steps, interpolate_fn, unipc_bh_rhos,  # noqa: F401
                                    unipc_multistep_table)
from diffvits_b200.schedule import model_wrapper as _model_wrapper

__all__ = ["NoiseScheduleVP", "model_wrapper", "UniPC", "interpolate_fn", "expand_dims"]


class NoiseScheduleVP(_NoiseScheduleVP):
    def __init__(self, schedule="discrete", betas=None, alphas_cumprod=None, continuous_beta_0=0.1,
                 continuous_beta_1=20.0, dtype=torch.float32):
        super().__init__(schedule, betas, alphas_cumprod, continuous_beta_0, continuous_beta_1, dtype,
                         clip_lambda=None, allow_cosine=True)


def model_wrapper(model, noise_schedule, model_type="noise", model_kwargs={}, guidance_type="uncond", condition=None,
                  unconditional_condition=None, guidance_scale=1.0, classifier_fn=None, classifier_kwargs={}):
    return _model_wrapper(model, noise_schedule, model_type, model_kwargs, guidance_type, condition,
                          unconditional_condition, guidance_scale, classifier_fn, classifier_kwargs,
                          batched_coeffs=True)


class UniPC:
    def __init__(self, model_fn, noise_schedule, algorithm_type="data_prediction", correcting_x0_fn=None,
                 correcting_xt_fn=None, thresholding_max_val=1.0, dynamic_thresholding_ratio=0.995, variant="bh1"):
        assert algorithm_type in ["data_prediction", "noise_prediction"]
        self._wrapped = model_fn
        self.model = lambda x, t: model_fn(x, t.expand((x.shape[0])))
        self.noise_schedule = noise_schedule
        self.correcting_x0_fn = (self.dynamic_thresholding_fn if correcting_x0_fn == "dynamic_thresholding"
                                 else correcting_x0_fn)
        self.correcting_xt_fn = correcting_xt_fn
        self.dynamic_thresholding_ratio = dynamic_thresholding_ratio
        self.thresholding_max_val = thresholding_max_val
        self.variant = variant
        self.predict_x0 = algorithm_type == "data_prediction"
        self.last_run_native = False

    def dynamic_thresholding_fn(self, x0, t=None):
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
            x0 = self.correcting_x0_fn(x0)
        return x0

    def model_fn(self, x, t):
        return self.data_prediction_fn(x, t) if self.predict_x0 else self.noise_prediction_fn(x, t)

    def get_time_steps(self, skip_type, t_T, t_0, N, device):
        return get_time_steps(self.noise_schedule, skip_type, t_T, t_0, N, device)

    def denoise_to_zero_fn(self, x, s):
        return self.data_prediction_fn(x, s)

    # ------------------------------------------------------------------ one predictor(-corrector) step
    def multistep_uni_pc_update(self, x, model_prev_list, t_prev_list, t, order, **kwargs):
        if len(t.shape) == 0:
            t = t.view(-1)
        if "bh" in self.variant:
            return self.multistep_uni_pc_bh_update(x, model_prev_list, t_prev_list, t, order, **kwargs)
        assert self.variant == "vary_coeff"
        return self.multistep_uni_pc_vary_update(x, model_prev_list, t_prev_list, t, order, **kwargs)

    def multistep_uni_pc_vary_update(self, x, model_prev_list, t_prev_list, t, order, use_corrector=True):
        """UniPC with varying coefficients (uni_pc.py:368-469). With r_i = (lambda_{-i} - lambda_0) / h for the history
        points (r_K = 1 for the new point), D_i = (m_{-i} - m_0) / r_i and the Vandermonde-like matrix
        C[i, k] = r_i^k / (k+1)! , the predictor adds  -A * sum_k phi_{k+1}(h) * (C_p^{-1}[k] . D)  to the first-order step
        and the corrector repeats that with the full C^{-1} and the new difference (m_t - m_0). Returns (x_t, model_t)."""
        ns = self.noise_schedule
        assert order <= len(model_prev_list)
        t0, m0 = t_prev_list[-1], model_prev_list[-1]
        lam0 = ns.marginal_lambda(t0)
        h = ns.marginal_lambda(t) - lam0
        ratios, diffs = [], []
        for i in range(1, order):
            r = (ns.marginal_lambda(t_prev_list[-(i + 1)]) - lam0) / h
            ratios.append(r)
            diffs.append((model_prev_list[-(i + 1)] - m0) / r)
        ratios.append(1.0)
        rks = torch.tensor(ratios, device=x.device)
        K = rks.numel()
        cols, col = [], torch.ones_like(rks)
        for k in range(1, K + 1):            # column k-1 holds r^(k-1) / k!
            cols.append(col)
            col = col * rks / (k + 1)
        Cm = torch.stack(cols, dim=1)
        inv_p = torch.linalg.inv(Cm[:-1, :-1]) if diffs else None
        inv_c = torch.linalg.inv(Cm) if use_corrector else None
        hh = -h if self.predict_x0 else h
        # phi_1(h) h, then the recurrence phi_{k+1} = phi_k / h - 1/k!
        phis, phi, fact = [], torch.expm1(hh), 1
        for k in range(1, K + 2):
            phis.append(phi)
            phi = phi / hh - 1 / fact
            fact *= (k + 1)
        if self.predict_x0:
            scale = ns.marginal_alpha(t)
            lin = ns.marginal_std(t) / ns.marginal_std(t0) * x - scale * phis[0] * m0
        else:
            scale = ns.marginal_std(t)
            lin = torch.exp(ns.marginal_log_mean_coeff(t) - ns.marginal_log_mean_coeff(t0)) * x - (scale * phis[0]) * m0
        D = torch.stack(diffs, dim=1) if diffs else None          # [B, K-1, ...]

        def mix(weights):                                         # sum_j weights[j] * D[:, j]
            return torch.einsum("bk...,k->b...", D, weights)

        x_t = lin
        for k in range(K - 1):
            x_t = x_t - scale * phis[k + 1] * mix(inv_p[k])
        model_t = None
        if use_corrector:
            model_t = self.model_fn(x_t, t)
            x_t = lin
            for k in range(K - 1):
                x_t = x_t - scale * phis[k + 1] * mix(inv_c[k][:-1])
            last_row = max(K - 2, 0)     # the reference weighs the new difference with row K-2 of C^{-1} (row 0 when K = 1)
            x_t = x_t - scale * phis[K] * ((model_t - m0) * inv_c[last_row][-1])
        return x_t, model_t

    def multistep_uni_pc_bh_update(self, x, model_prev_list, t_prev_list, t, order, x_t=None, use_corrector=True):
        """UniPC-B(h) step of the given order from t_prev_list[-1] to t (uni_pc.py:471-588). Returns (x_t, model_t);
        model_t is the model evaluated at the PREDICTED point and is reused as the next history entry."""
        ns = self.noise_schedule
        assert order <= len(model_prev_list)
        t0, m0 = t_prev_list[-1], model_prev_list[-1]
        lam0, lam_t = ns.marginal_lambda(t0), ns.marginal_lambda(t)
        h = lam_t - lam0
        rks, D1s = [], []
        for i in range(1, order):
            lam_i = ns.marginal_lambda(t_prev_list[-(i + 1)])
            rk = ((lam_i - lam0) / h)[0]
            rks.append(rk)
            D1s.append((model_prev_list[-(i + 1)] - m0) / rk)
        rks.append(torch.ones((), device=x.device, dtype=h.dtype))
        hh = (-h if self.predict_x0 else h)[0]
        h_phi_1, b_h, rhos_p, rhos_c = unipc_bh_rhos(hh.cpu(), [r.cpu() for r in rks], self.variant)
        h_phi_1, b_h = h_phi_1.to(x.device), b_h.to(x.device)
        if self.predict_x0:
            lin = ns.marginal_std(t) / ns.marginal_std(t0) * x - ns.marginal_alpha(t) * h_phi_1 * m0
            gain = ns.marginal_alpha(t) * b_h
        else:
            la0, la_t = ns.marginal_log_mean_coeff(t0), ns.marginal_log_mean_coeff(t)
            lin = torch.exp(la_t - la0) * x - ns.marginal_std(t) * h_phi_1 * m0
            gain = ns.marginal_std(t) * b_h
        if x_t is None:
            pred = sum(float(rhos_p[k]) * D1s[k] for k in range(len(D1s))) if D1s else 0
            x_t = lin - gain * pred
        model_t = None
        if use_corrector:
            model_t = self.model_fn(x_t, t)
            corr = sum(float(rhos_c[k]) * D1s[k] for k in range(len(D1s))) if D1s else 0
            x_t = lin - gain * (corr + float(rhos_c[-1]) * (model_t - m0))
        return x_t, model_t

    # ------------------------------------------------------------------ native fast path
    def _native(self, method, order, denoise_to_zero, return_intermediate):
        dv = getattr(self._wrapped, "_dv", None)
        if (dv is None or not self.predict_x0 or method != "multistep" or order > 2
                or self.variant not in ("bh1", "bh2") or denoise_to_zero or return_intermediate
                or self.correcting_x0_fn is not None or self.correcting_xt_fn is not None
                or dv["model_type"] != "x_start" or dv["guidance_type"] != "uncond"
                or dv["noise_schedule"] is not self.noise_schedule):
            return None
        from diffvits_b200.native import find_native
        return find_native(dv["model"], dv["model_kwargs"])

    # ------------------------------------------------------------------ sample
    def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type="time_uniform", method="multistep",
               lower_order_final=True, denoise_to_zero=False, atol=0.0078, rtol=0.05, return_intermediate=False):
        ns = self.noise_schedule
        t_0 = 1.0 / ns.total_N if t_end is None else t_end
        t_T = ns.T if t_start is None else t_start
        assert t_0 > 0 and t_T > 0, ("Time range needs to be greater than 0. For discrete-time DPMs, it needs to be "
                                     "in [1 / N, 1], where N is the length of betas array")
        if method != "multistep":
            raise ValueError("Got wrong method {}".format(method))
        assert steps >= order
        self.last_run_native = False
        native = self._native(method, order, denoise_to_zero, return_intermediate)
        if native is not None and x.is_cuda:
            from diffvits_b200.engine import SAMPLER_UNIPC_BH2
            from diffvits_b200.native import tick_progress
            ts = self.get_time_steps(skip_type, t_T, t_0, steps, "cpu")
            t_in, coef = unipc_multistep_table(ns, ts, order, lower_order_final, self.variant)
            out = native.engine.sample(SAMPLER_UNIPC_BH2, t_in, coef, x.to(torch.float32).contiguous().clone())
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
                    step_order, use_corrector = step, True
                else:
                    step_order = min(order, steps + 1 - step) if lower_order_final else order
                    use_corrector = step < steps            # no corrector (and no model call) on the last step
                x, model_x = self.multistep_uni_pc_update(x, m_hist, t_hist, t, step_order,
                                                          use_corrector=use_corrector)
                if self.correcting_xt_fn is not None:
                    x = self.correcting_xt_fn(x, t, step)
                if return_intermediate:
                    intermediates.append(x)
                if step < steps:
                    if model_x is None:
                        model_x = self.model_fn(x, t)
                    t_hist.append(t)
                    m_hist.append(model_x)
                    del t_hist[:-order], m_hist[:-order]
            if denoise_to_zero:
                t = torch.ones((1,)).to(device) * t_0
                x = self.denoise_to_zero_fn(x, t)
                if self.correcting_xt_fn is not None:
                    x = self.correcting_xt_fn(x, t, steps + 1)
                if return_intermediate:
                    intermediates.append(x)
        return (x, intermediates) if return_intermediate else x
