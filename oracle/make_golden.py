"""ORACLE tooling: generate tests/golden/* by running the REFERENCE (imported from /root/reference) on seeded
weights and inputs. Run in the build container: `python -m oracle.make_golden`. The outputs are small and committed;
the script is the provenance record for them."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import inputs, ref_import  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


@torch.no_grad()
def main():
    assert ref_import.available(), "reference checkout not found"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    unet = ref_import.build_ref_unet(1234)
    sd = unet.state_dict()

    # 1. weight checksums (proves that a module built elsewhere from seed 1234 has the reference's weights)
    chk = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}
    with open(os.path.join(GOLD, "weights_seed1234.json"), "w") as f:
        json.dump({"n_tensors": len(chk), "n_params": int(sum(v.numel() for v in sd.values())), "sums": chk}, f)

    # 2. single denoiser calls, C1 shapes (B=1, T=200, Tp=282) at three model times
    x_T, cond, ehs = inputs.make_inputs(1, 200, 282)
    sample = torch.cat([x_T, cond], dim=1)
    ts = [999.0, 499.5, 0.0]
    outs = [unet(sample, torch.tensor([t]), ehs).sample.numpy() for t in ts]
    np.savez_compressed(os.path.join(GOLD, "unet_c1.npz"), t=np.array(ts, np.float32), out=np.stack(outs))

    # 3. ragged prompts: B=2, T=64, lengths [282, 182], masked rows zeroed (model3.py:905-912)
    x2, c2, e2 = inputs.make_inputs(2, 64, 282, seed=1)
    lens, mask = inputs.ragged_mask(2, 282)
    e2 = e2 * mask[:, :, None]
    out2 = unet(torch.cat([x2, c2], 1), torch.tensor([250.25, 250.25]), e2, encoder_attention_mask=mask).sample
    np.savez_compressed(os.path.join(GOLD, "unet_ragged.npz"), lens=lens.numpy(), out=out2.numpy())

    # 3b. lengths that are not multiples of 8 (203 -> 102 -> 51 -> 26 frames: the stride-2 convs see odd inputs, the
    # nearest resizes go to 2L - 1 = 51, 2L = 102 and 2L - 1 = 203 rows; 198 -> 99 -> 50 -> 25 the other way round)
    odd = {}
    for T_odd, seed in ((203, 3), (198, 4)):
        x3, c3, e3 = inputs.make_inputs(1, T_odd, 40, seed=seed)
        odd[f"out{T_odd}"] = unet(torch.cat([x3, c3], 1), torch.tensor([613.0]), e3).sample.numpy()
    np.savez_compressed(os.path.join(GOLD, "unet_oddlen.npz"), **odd)

    # 4. full sampling loops at C1 with the reference samplers wired like model3.py:1143-1158 / 1177-1192
    rdpm, runi = ref_import.load_samplers()
    betas = inputs.betas_linear()
    step_outs, step_x, step_t = [], [], []

    def model(x, t, **kw):
        out = unet(torch.cat([x, cond], dim=1), t, ehs).sample
        step_outs.append(out)
        step_x.append(x.clone())          # the denoiser INPUT of every call: teacher-forced per-step parity needs x_k
        step_t.append(t.clone())
        return out

    ns = rdpm.NoiseScheduleVP("discrete", betas=betas)
    fn = rdpm.model_wrapper(model, ns, model_type="x_start", model_kwargs={})
    lat_dpm = rdpm.DPM_Solver(fn, ns, algorithm_type="dpmsolver++").sample(
        x_T.clone(), steps=50, order=2, skip_type="time_uniform", method="multistep")
    dpm_steps = torch.stack([step_outs[i] for i in (0, 1, 25, 49)]).numpy()
    dpm_x = torch.stack([step_x[i] for i in (0, 1, 25, 49)]).numpy()
    dpm_t = torch.stack([step_t[i].reshape(-1)[0] for i in (0, 1, 25, 49)]).numpy()
    assert len(step_outs) == 50
    step_outs.clear(); step_x.clear(); step_t.clear()
    ns2 = runi.NoiseScheduleVP("discrete", betas=betas)
    fn2 = runi.model_wrapper(model, ns2, model_type="x_start", model_kwargs={})
    lat_uni = runi.UniPC(fn2, ns2, variant="bh2").sample(
        x_T.clone(), steps=10, order=2, skip_type="time_uniform", method="multistep")
    assert len(step_outs) == 10
    uni_steps = torch.stack([step_outs[i] for i in (0, 1, 9)]).numpy()
    uni_x = torch.stack([step_x[i] for i in (0, 1, 9)]).numpy()
    uni_t = torch.stack([step_t[i].reshape(-1)[0] for i in (0, 1, 9)]).numpy()
    np.savez_compressed(os.path.join(GOLD, "sampler_c1.npz"), dpm50=lat_dpm.numpy(), dpm_steps=dpm_steps,
                        unipc10=lat_uni.numpy(), uni_steps=uni_steps)
    # denoiser inputs (x_k, model time) at the same trajectory points, fp16-free fp32: ~0.6 MB
    np.savez_compressed(os.path.join(GOLD, "sampler_c1_inputs.npz"), dpm_x=dpm_x, dpm_t=dpm_t, uni_x=uni_x, uni_t=uni_t)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
