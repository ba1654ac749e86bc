"""Native PromptEncoder (csrc/prompt_encoder.cu) vs the reference's own module (model3.PromptEncoder from the oracle/_ref
snapshot, fp32 eager on the same GPU), ragged prompt lengths. Prints one JSON line. Run by tests/test_gpu_reference.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_snapshot  # noqa: E402

ref_snapshot.activate("dropin")
import torch  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
model3 = ref_snapshot.import_model3()
from diffvits_b200 import prompt as native_prompt  # noqa: E402

dev = "cuda:0"
torch.manual_seed(4321)
pe = model3.PromptEncoder(100, 128, 128, 4, 0.2).eval().to(dev)
# the reference initialises LayerNorm affine to (1, 0) and conv biases to 0: perturb them so every parameter is exercised
with torch.no_grad():
    for k, v in pe.named_parameters():
        if k.endswith("bias") or "layer_norm" in k:
            v.add_(0.1 * torch.randn_like(v))
assert native_prompt.supported(pe) is not None
res = {}
for name, B, Tp, lens in (("full", 2, 282, [282, 282]), ("ragged", 4, 200, [200, 131, 64, 3]), ("short", 1, 17, [17])):
    g = torch.Generator().manual_seed(5)
    prompt = torch.randn(B, 100, Tp, generator=g).to(dev)
    pl = torch.tensor(lens, device=dev)
    with torch.no_grad():
        mask = (torch.arange(Tp, device=dev)[None, :] < pl[:, None])
        ref = (pe(prompt, pl) * mask.unsqueeze(1).to(prompt.dtype)).transpose(1, 2).contiguous()
        got = native_prompt.encode_prompt(pe, prompt, pl)
    d = got.double() - ref.double()
    pad_zero = bool((got[~mask] == 0).all())
    res[name] = {"rel_l2": float(d.norm() / ref.double().norm()), "max_abs": float(d.abs().max()), "pad_rows_zero": pad_zero,
                 "finite": bool(torch.isfinite(got).all())}
print(json.dumps(res))
