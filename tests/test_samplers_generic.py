"""Sampler settings outside the native fast path on the drop-in `sampler.dpm_solver`:
singlestep ("DPM-Solver-fast"), singlestep_fixed and adaptive solvers in both parameterisations, checked against
golden outputs of the REFERENCE (tests/golden/sampler_generic.npz, oracle/make_golden_samplers.py) and, when the
reference checkout is present, against the live reference over a wider sweep. fp32 tolerance: 1e-5 relative."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import ref_import
from oracle.make_golden_samplers import CASES, UNIPC_CASES, run_case, run_unipc_case, toy_inputs


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def test_generic_solvers_vs_reference_golden(dropin):
    g = np.load(os.path.join(GOLD, "sampler_generic.npz"))["out"]
    x, betas = toy_inputs()
    assert len(g) == len(CASES)
    for case, want in zip(CASES, g):
        got = run_case(dropin["dpm"], case, x, betas)
        assert torch.isfinite(got).all(), case
        assert _rel(got, torch.from_numpy(want)) < 1e-5, case


def test_unipc_variants_vs_reference_golden(dropin):
    g = np.load(os.path.join(GOLD, "sampler_generic.npz"))["unipc"]
    x, betas = toy_inputs()
    assert len(g) == len(UNIPC_CASES)
    for case, want in zip(UNIPC_CASES, g):
        got = run_unipc_case(dropin["uni"], case, x[:1], betas)
        assert _rel(got, torch.from_numpy(want)) < 1e-5, case
    # batch > 1 works in the drop-in (the reference raises for x_start models) and equals the per-utterance runs
    case = UNIPC_CASES[0]
    both = run_unipc_case(dropin["uni"], case, x, betas)
    for b in range(x.shape[0]):
        assert _rel(both[b:b + 1], run_unipc_case(dropin["uni"], case, x[b:b + 1], betas)) < 1e-6


def test_singlestep_order_schedule(dropin):
    """NFE bookkeeping of DPM-Solver-fast: the orders add up to `steps` and the outer grid has one point per step."""
    dpm = dropin["dpm"]
    ns = dpm.NoiseScheduleVP("discrete", betas=torch.linspace(1e-4, 2e-2, 1000))
    solver = dpm.DPM_Solver(lambda x, t: x, ns)
    for steps in range(3, 14):
        for order in (1, 2, 3):
            for skip in ("time_uniform", "logSNR"):
                if order == 1 and skip == "logSNR":
                    continue                      # the reference raises IndexError for this combination
                outer, orders = solver.get_orders_and_timesteps_for_singlestep_solver(steps, order, skip, 1.0, 1e-3, "cpu")
                assert sum(orders) == steps and max(orders) <= order
                assert len(outer) == len(orders) + 1
                assert float(outer[0]) == pytest.approx(1.0) and float(outer[-1]) == pytest.approx(1e-3, rel=1e-4)
    with pytest.raises(ValueError):
        solver.sample(torch.zeros(1, 2, 4), steps=4, method="nope")


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present")
def test_generic_solvers_vs_live_reference(dropin):
    rdpm, _ = ref_import.load_samplers()
    x, betas = toy_inputs()
    n = 0
    for alg in ("dpmsolver++", "dpmsolver"):
        for mt in ("noise", "x_start"):
            for method, steps, order in (("singlestep", 10, 3), ("singlestep", 11, 3), ("singlestep", 9, 2),
                                         ("singlestep_fixed", 12, 3), ("adaptive", 0, 2), ("adaptive", 0, 3)):
                for st in ("dpmsolver", "taylor"):
                    for skip in ("time_uniform", "logSNR", "time_quadratic"):
                        if method == "adaptive" and skip != "time_uniform":
                            continue
                        case = (alg, mt, method, steps, order, st, skip, steps % 2 == 0)
                        assert _rel(run_case(dropin["dpm"], case, x, betas), run_case(rdpm, case, x, betas)) < 1e-5, case
                        n += 1
    assert n == 112
    _, runi = ref_import.load_samplers()
    for variant in ("vary_coeff", "bh1", "bh2"):
        for alg in ("data_prediction", "noise_prediction"):
            for mt in ("noise", "x_start"):
                for steps, order in ((10, 1), (10, 2), (12, 3), (5, 2)):
                    for lof in (True, False):
                        case = (variant, alg, mt, steps, order, "time_uniform", lof)
                        a = run_unipc_case(dropin["uni"], case, x[:1], betas)
                        assert _rel(a, run_unipc_case(runi, case, x[:1], betas)) < 1e-5, case
