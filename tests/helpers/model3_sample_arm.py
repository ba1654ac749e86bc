"""One arm of the `model3.NaturalSpeech2.sample` drop-in check (tests/test_gpu_reference.py, run through
oracle.ref_snapshot.run_arm): builds the reference's NaturalSpeech2 from config.json with seeded random-init weights,
runs the UNMODIFIED `sample()` body (model3.py:1120-1203, `unipc` or `dpmsolver` branch... the latter is un-runnable in the
reference) with a pass-through vocoder, and saves the mel latent.

    DV_ARM=reference|dropin python model3_sample_arm.py OUT.npz [device] [n_text] [Tp] [seed]

DV_ARM=reference: every class is the reference's.  DV_ARM=dropin: `unet1d` / `sampler` are the drop-in packages, so the
diffusion denoiser + solver loop run in libdiffvits_b200 (find_native case 2: closure unwrap, PromptEncoder once).
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_snapshot  # noqa: E402

arm = os.environ.get("DV_ARM", "reference")
out_path = sys.argv[1]
device = sys.argv[2] if len(sys.argv) > 2 else "cpu"
n_text = int(sys.argv[3]) if len(sys.argv) > 3 else 24
Tp = int(sys.argv[4]) if len(sys.argv) > 4 else 120
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 0
length_scale = float(os.environ.get("DV_LENGTH_SCALE", "1.0"))

ref_snapshot.activate(arm)
import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
model3 = ref_snapshot.import_model3()
from text import symbols  # noqa: E402  (the reference's symbol table)

cfg = ref_snapshot.load_config()
torch.manual_seed(1234)
model = model3.NaturalSpeech2(cfg).eval()
unet_mod = type(model.diff_model.unet).__module__
model.to(device)


class PassThroughVocoder:
    """Stands in for Vocos (absent here; outside the hot path). The mel latent is what is compared."""

    def to(self, *_a, **_k):
        return self

    def decode(self, mel):
        return mel.new_zeros(mel.shape[0], 256 * mel.shape[-1])      # [B, samples]: silence of the right length


g = torch.Generator().manual_seed(seed)
text = torch.randint(0, len(symbols), (1, n_text), generator=g).to(device)
tone = torch.randint(0, 4, (1, n_text), generator=g).to(device)
language = torch.zeros(1, n_text, dtype=torch.long).to(device)
refer = torch.randn(1, 100, Tp, generator=g).to(device)
text_len = torch.tensor([n_text]).to(device)
refer_len = torch.tensor([Tp]).to(device)

if length_scale != 1.0:
    # random-init duration predictor: scale the durations so the utterance has a useful number of frames (the
    # reference exposes `length_scale` on VITS.infer, model3.py:826; sample() does not forward it)
    _infer = model.vits.infer
    model.vits.infer = lambda *a, **k: _infer(*a, length_scale=length_scale, **k)

torch.manual_seed(seed + 1)           # vits.infer noise (z_p) and the initial latent (model3.py:1176) come from here
if device != "cpu":
    torch.cuda.manual_seed(seed + 1)
t0 = time.perf_counter()
with torch.no_grad():
    audio, mel = model.sample(text, refer, text_len, refer_len, tone, language, PassThroughVocoder())
if device != "cpu":
    torch.cuda.synchronize()
dt = time.perf_counter() - t0
native = None
if arm == "dropin":
    import diffvits_b200.native as _n
    native = _n.NATIVE_LOOPS
np.savez_compressed(out_path, mel=mel.float().cpu().numpy(), seconds=dt)
print(f"arm={arm} unet={unet_mod} mel={tuple(mel.shape)} finite={bool(torch.isfinite(mel).all())} "
      f"seconds={dt:.2f} native_loop={native}")
