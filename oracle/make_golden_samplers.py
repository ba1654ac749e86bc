"""ORACLE tooling: golden vectors for the sampler settings that run the generic Python loops (SURVEY.md section 8,
row a21): singlestep / singlestep_fixed / adaptive DPM-Solver in both parameterisations and solver types, and the UniPC
variants (vary_coeff, bh1, bh2 at orders 1-3), produced by the REFERENCE's `sampler/dpm_solver.py` / `sampler/uni_pc.py`
(imported from /root/reference) on a small analytic denoiser.
Run in the build container: `python -m oracle.make_golden_samplers`; writes tests/golden/sampler_generic.npz."""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (algorithm_type, model_type, method, steps, order, solver_type, skip_type, denoise_to_zero)
CASES = [
    ("dpmsolver++", "x_start", "singlestep", 10, 3, "dpmsolver", "time_uniform", False),
    ("dpmsolver++", "x_start", "singlestep", 11, 3, "taylor", "logSNR", True),
    ("dpmsolver++", "noise", "singlestep", 12, 3, "dpmsolver", "time_quadratic", False),
    ("dpmsolver++", "x_start", "singlestep", 7, 2, "dpmsolver", "time_uniform", False),
    ("dpmsolver", "noise", "singlestep", 8, 2, "taylor", "logSNR", False),
    ("dpmsolver", "noise", "singlestep", 10, 3, "dpmsolver", "time_uniform", True),
    ("dpmsolver", "x_start", "singlestep", 5, 1, "dpmsolver", "time_uniform", False),
    ("dpmsolver++", "x_start", "singlestep_fixed", 12, 3, "dpmsolver", "time_uniform", False),
    ("dpmsolver", "noise", "singlestep_fixed", 12, 2, "taylor", "time_uniform", False),
    ("dpmsolver++", "x_start", "adaptive", 0, 2, "dpmsolver", "time_uniform", False),
    ("dpmsolver++", "x_start", "adaptive", 0, 3, "taylor", "time_uniform", False),
    ("dpmsolver", "noise", "adaptive", 0, 3, "dpmsolver", "time_uniform", False),
]


# UniPC (variant, algorithm_type, model_type, steps, order, skip_type, lower_order_final); batch 1 because the reference's
# uni_pc.model_wrapper cannot broadcast alpha_t / sigma_t over a batch for x_start models (uni_pc.py:189-191)
UNIPC_CASES = [
    ("vary_coeff", "data_prediction", "x_start", 10, 2, "time_uniform", True),
    ("vary_coeff", "data_prediction", "x_start", 12, 3, "logSNR", True),
    ("vary_coeff", "noise_prediction", "noise", 10, 3, "time_uniform", False),
    ("vary_coeff", "data_prediction", "noise", 6, 1, "time_uniform", True),
    ("bh1", "data_prediction", "x_start", 10, 2, "time_uniform", True),
    ("bh1", "noise_prediction", "noise", 12, 3, "logSNR", True),
    ("bh2", "data_prediction", "x_start", 12, 3, "time_uniform", True),
    ("bh2", "noise_prediction", "x_start", 5, 2, "time_uniform", False),
]


def toy_model(x, t, **kw):
    """Smooth, nonlinear, time-dependent stand-in for the denoiser (t is the discrete model time in [0, 1000))."""
    return torch.tanh(x * 0.7) + 0.1 * torch.sin(t / 100.0).reshape(-1, 1, 1) * x


def toy_inputs():
    g = torch.Generator().manual_seed(5)
    return torch.randn(2, 5, 16, generator=g), torch.linspace(1e-4, 2e-2, 1000)


def run_case(mod, case, x, betas):
    alg, mt, method, steps, order, st, skip, dtz = case
    ns = mod.NoiseScheduleVP("discrete", betas=betas)
    fn = mod.model_wrapper(toy_model, ns, model_type=mt, model_kwargs={})
    solver = mod.DPM_Solver(fn, ns, algorithm_type=alg)
    with contextlib.redirect_stdout(io.StringIO()):      # the reference prints the adaptive solver's NFE
        return solver.sample(x.clone(), steps=steps, order=order, skip_type=skip, method=method, solver_type=st,
                             denoise_to_zero=dtz)


def run_unipc_case(mod, case, x, betas):
    variant, alg, mt, steps, order, skip, lof = case
    ns = mod.NoiseScheduleVP("discrete", betas=betas)
    fn = mod.model_wrapper(toy_model, ns, model_type=mt, model_kwargs={})
    return mod.UniPC(fn, ns, algorithm_type=alg, variant=variant).sample(
        x.clone(), steps=steps, order=order, skip_type=skip, method="multistep", lower_order_final=lof)


@torch.no_grad()
def main():
    assert ref_import.available(), "reference checkout not found"
    rdpm, runi = ref_import.load_samplers()
    x, betas = toy_inputs()
    outs = np.stack([run_case(rdpm, c, x, betas).numpy() for c in CASES])
    uni = np.stack([run_unipc_case(runi, c, x[:1], betas).numpy() for c in UNIPC_CASES])
    np.savez_compressed(os.path.join(GOLD, "sampler_generic.npz"), out=outs, unipc=uni)
    print("sampler_generic.npz", outs.shape, uni.shape)


if __name__ == "__main__":
    main()
