"""`Diffusion_Encoder.forward` (model3.py:902-914) over the drop-in UNet vs over the reference UNet, same weights, same
inputs, in one process on one CUDA device. Prints one JSON line of distances. Run by tests/test_gpu_reference.py.

The module is built by the reference's own model3.Diffusion_Encoder with `unet1d` resolving to the drop-in package; the
reference's UNet class is then imported under a private name (oracle.ref_import), given the same state_dict and swapped
in for the second evaluation. Ragged prompt lengths exercise the mask -> kv_len adapter of the Tier-A call."""
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
from oracle import ref_import  # noqa: E402

cfg = ref_snapshot.load_config()
dev = "cuda:0"
torch.manual_seed(1234)
diff = model3.Diffusion_Encoder(**cfg["diffusion_encoder"]).eval().to(dev)
assert type(diff.unet).__module__ == "unet1d.unet_1d_condition" and "dropin" in sys.modules["unet1d"].__file__
ref_unet = ref_import.load_unet_module().UNet1DConditionModel(**ref_import.UNET_KW).eval()
ref_unet.load_state_dict(diff.unet.state_dict())
ref_unet.to(dev)

res = {}
for name, B, T, Tp, lens in (("full", 2, 200, 282, [282, 282]), ("ragged", 3, 136, 282, [282, 150, 1])):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 100, T, generator=g).to(dev)
    cond = torch.randn(B, 128, T, generator=g).to(dev)
    prompt = torch.randn(B, 100, Tp, generator=g).to(dev)
    pl = torch.tensor(lens, device=dev)
    cl = torch.full((B,), T, device=dev)
    t = torch.full((B,), 417.25, device=dev)
    data = (cond, prompt, cl, pl)
    with torch.no_grad():
        native_unet = diff.unet
        out_native = diff(x, data, t).clone()
        diff.unet = ref_unet
        out_ref = diff(x, data, t)
        diff.unet = native_unet
    d = (out_native.double() - out_ref.double())
    res[name] = {"rel_l2": float(d.norm() / out_ref.double().norm()), "max_abs": float(d.abs().max()),
                 "finite": bool(torch.isfinite(out_native).all())}
print(json.dumps(res))
