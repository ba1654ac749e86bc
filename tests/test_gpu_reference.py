"""The REAL reference on the GPU box (snapshot oracle/_ref, made by oracle/make_ref_snapshot.py): the reference's own call
sites run over the drop-in packages and over the reference's classes, and the results are compared.

  * `model3.NaturalSpeech2.sample` (model3.py:1120-1203, `unipc` branch, 30 steps) unmodified, with a
    pass-through vocoder. Drop-in arm: `find_native` case 2 (closure unwrap, PromptEncoder hoisted to once per job) must
    reach the native whole-loop path; the mel latent is compared with the all-reference arm (fp32 eager on the same GPU).
  * row a2: `Diffusion_Encoder.forward` (model3.py:902-914): bool mask -> kv_len adapter, ragged prompt lengths.
Each arm is its own process: the reference resolves `unet1d` / `sampler` by module name."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
HELPERS = os.path.join(ROOT, "tests", "helpers")


@pytest.fixture(scope="module")
def snap():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from oracle import ref_snapshot
    if not ref_snapshot.available():
        pytest.fail("oracle/_ref snapshot missing: run `python -m oracle.make_ref_snapshot` in the build container")
    return ref_snapshot


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))


def test_model3_sample_over_dropin_matches_reference_classes(snap, tmp_path):
    outs = {}
    for arm in ("reference", "dropin"):
        path = str(tmp_path / f"{arm}.npz")
        r = snap.run_arm(os.path.join(HELPERS, "model3_sample_arm.py"), arm, [path, "cuda:0", 24, 120, 0],
                         env={"DV_LENGTH_SCALE": "6"})
        assert r.returncode == 0, r.stderr[-3000:]
        last = r.stdout.strip().splitlines()[-1]
        print(last)
        outs[arm] = (np.load(path), last)
    ref, dro = outs["reference"][0]["mel"], outs["dropin"][0]["mel"]
    assert "unet=unet1d.unet_1d_condition" in outs["dropin"][1] and "native_loop=1" in outs["dropin"][1], outs["dropin"][1]
    assert ref.shape == dro.shape and ref.shape[1] == 100 and ref.shape[2] >= 100, ref.shape
    assert np.isfinite(dro).all()
    rel, mx = _rel(dro, ref), float(np.abs(dro - ref).max())
    print(f"NaturalSpeech2.sample (unipc, 30 steps, T={ref.shape[2]}): drop-in vs reference classes rel-L2 {rel:.3e} max-abs {mx:.3e}; "
          f"seconds reference {float(outs['reference'][0]['seconds']):.2f} drop-in {float(outs['dropin'][0]['seconds']):.2f}")
    assert rel < 4e-2            # BF16 operands, free-running 30-step loop (floor measured in test_gpu_unet.py)


def test_diffusion_encoder_forward_over_dropin_matches_reference_unet(snap):
    r = snap.run_arm(os.path.join(HELPERS, "diffenc_check.py"), "dropin")
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    print(res)
    for name, m in res.items():
        assert m["finite"], name
        assert m["rel_l2"] < 1.5e-2 and m["max_abs"] < 2e-2, (name, m)


def test_native_prompt_encoder_matches_reference_module(snap):
    """PromptEncoder (model3.py:382-433) on the engine's kernels vs the reference module, ragged lengths. BF16
    operands through 4 transformer layers: the same per-call floor as the UNet (measured ~4e-3)."""
    r = snap.run_arm(os.path.join(HELPERS, "prompt_encoder_check.py"), "dropin")
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    print(res)
    for name, m in res.items():
        assert m["finite"] and m["pad_rows_zero"], (name, m)
        assert m["rel_l2"] < 1e-2 and m["max_abs"] < 5e-2, (name, m)
