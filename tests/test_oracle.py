"""The oracle (oracle/*.py) is pinned against outputs of the reference itself: committed golden vectors
(tests/golden, made by oracle/make_golden.py) and, when /root/reference is present, the live reference."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import inputs, ref_import, sampler_oracle, unet_oracle


def rel_l2(a, b):
    return float((a - b).norm() / b.norm())


def test_unet_oracle_matches_golden_c1(seeded_unet):
    g = np.load(os.path.join(GOLD, "unet_c1.npz"))
    x_T, cond, ehs = inputs.make_inputs(1, 200, 282)
    sample = torch.cat([x_T, cond], 1)
    sd = seeded_unet.state_dict()
    for t, ref in zip(g["t"], g["out"]):
        out = unet_oracle.unet_forward(sd, sample, torch.tensor([float(t)]), ehs)
        ref = torch.from_numpy(ref)
        assert rel_l2(out, ref) < 2e-5 and float((out - ref).abs().max()) < 1e-4   # fp32 gate: 1e-4


def test_unet_oracle_matches_golden_ragged(seeded_unet):
    g = np.load(os.path.join(GOLD, "unet_ragged.npz"))
    x, c, e = inputs.make_inputs(2, 64, 282, seed=1)
    lens, mask = inputs.ragged_mask(2, 282)
    assert lens.tolist() == g["lens"].tolist()
    e = e * mask[:, :, None]
    out = unet_oracle.unet_forward(seeded_unet.state_dict(), torch.cat([x, c], 1), torch.tensor([250.25, 250.25]), e,
                                   mask=mask)
    ref = torch.from_numpy(g["out"])
    assert rel_l2(out, ref) < 2e-5 and float((out - ref).abs().max()) < 1e-4


def test_unet_oracle_matches_golden_odd_lengths(seeded_unet):
    """Frame counts that are not multiples of 8: stride-2 convs over odd lengths, nearest resize to 2L - 1 rows."""
    g = np.load(os.path.join(GOLD, "unet_oddlen.npz"))
    sd = seeded_unet.state_dict()
    for T, seed in ((203, 3), (198, 4)):
        x, c, e = inputs.make_inputs(1, T, 40, seed=seed)
        ref = torch.from_numpy(g[f"out{T}"])
        out = unet_oracle.unet_forward(sd, torch.cat([x, c], 1), torch.tensor([613.0]), e)
        assert out.shape == ref.shape == (1, 100, T)
        assert rel_l2(out, ref) < 2e-5 and float((out - ref).abs().max()) < 1e-4
        emu = unet_oracle.unet_forward(sd, torch.cat([x, c], 1), torch.tensor([613.0]), e, emulate_bf16=True)
        assert rel_l2(emu, ref) < 1.5e-2


def test_bf16_emulation_is_close_to_fp32(seeded_unet):
    """BF16-operand emulation stays within the reference's own autocast-BF16 distance (SURVEY: 1.05e-2 rel-L2)."""
    x_T, cond, ehs = inputs.make_inputs(1, 64, 96)
    sd = seeded_unet.state_dict()
    sample = torch.cat([x_T, cond], 1)
    t = torch.tensor([500.0])
    a = unet_oracle.unet_forward(sd, sample, t, ehs)
    b = unet_oracle.unet_forward(sd, sample, t, ehs, emulate_bf16=True)
    assert rel_l2(b, a) < 2e-2


def _fake_model(x, t_in, **kw):
    return torch.tanh(x * 0.7) * torch.cos(t_in * 0.01).view(-1, 1, 1) + 0.1 * torch.sin(3 * x)


def test_sampler_oracle_matches_golden_c1(seeded_unet):
    """Free-running 50-step DPM-Solver++ / 10-step UniPC with the oracle UNet vs the reference's final latents."""
    g = np.load(os.path.join(GOLD, "sampler_c1.npz"))
    x_T, cond, ehs = inputs.make_inputs(1, 200, 282)
    sd = seeded_unet.state_dict()
    betas = inputs.betas_linear()
    torch.set_num_threads(os.cpu_count() or 1)

    def model(x, t_in):
        return unet_oracle.unet_forward(sd, torch.cat([x, cond], 1), t_in, ehs)

    outs = []
    lat = sampler_oracle.unipc_bh2(model, betas, x_T.clone(), 10, outs=outs)
    assert rel_l2(lat, torch.from_numpy(g["unipc10"])) < 1e-4
    for i, k in enumerate((0, 1, 9)):
        assert rel_l2(outs[k], torch.from_numpy(g["uni_steps"][i])) < 1e-4
    outs = []
    lat = sampler_oracle.dpmpp_2m(model, betas, x_T.clone(), 50, outs=outs)
    assert rel_l2(lat, torch.from_numpy(g["dpm50"])) < 1e-4
    for i, k in enumerate((0, 1, 25, 49)):
        assert rel_l2(outs[k], torch.from_numpy(g["dpm_steps"][i])) < 1e-4


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_sampler_oracle_matches_live_reference():
    rdpm, runi = ref_import.load_samplers()
    betas = inputs.betas_linear()
    g = torch.Generator().manual_seed(3)
    for B in (1, 3):
        x0 = torch.randn(B, 5, 24, generator=g)
        for steps in (8, 50):
            ns = rdpm.NoiseScheduleVP("discrete", betas=betas)
            fn = rdpm.model_wrapper(_fake_model, ns, model_type="x_start")
            ref = rdpm.DPM_Solver(fn, ns, "dpmsolver++").sample(x0.clone(), steps=steps, order=2,
                                                               skip_type="time_uniform", method="multistep")
            got = sampler_oracle.dpmpp_2m(_fake_model, betas, x0.clone(), steps)
            assert rel_l2(got, ref) < 2e-6
    x0 = torch.randn(1, 5, 24, generator=g)          # the reference's UniPC wrapper only works for batch 1
    for steps in (3, 10, 25):
        ns = runi.NoiseScheduleVP("discrete", betas=betas)
        fn = runi.model_wrapper(_fake_model, ns, model_type="x_start")
        ref = runi.UniPC(fn, ns, variant="bh2").sample(x0.clone(), steps=steps, order=2, skip_type="time_uniform",
                                                       method="multistep")
        got = sampler_oracle.unipc_bh2(_fake_model, betas, x0.clone(), steps)
        assert rel_l2(got, ref) < 2e-6


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_unet_oracle_matches_live_reference_small():
    unet = ref_import.build_ref_unet(1234)
    x_T, cond, ehs = inputs.make_inputs(2, 40, 50, seed=5)
    lens, mask = inputs.ragged_mask(2, 50, step=20)
    ehs = ehs * mask[:, :, None]
    t = torch.tensor([10.5, 900.0])
    with torch.no_grad():
        ref = unet(torch.cat([x_T, cond], 1), t, ehs, encoder_attention_mask=mask).sample
    out = unet_oracle.unet_forward(unet.state_dict(), torch.cat([x_T, cond], 1), t, ehs, mask=mask)
    assert rel_l2(out, ref) < 2e-5
