"""Small native forward + 3-step loop for compute-sanitizer (memcheck / racecheck): exercises the fused output GroupNorm, the
sampler diagnostics and the tensor-core split attention at tiny shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffvits_b200
from diffvits_b200 import synthetic as inputs
diffvits_b200.install_dropin()
from unet1d.unet_1d_condition import UNet1DConditionModel
torch.manual_seed(1234)
unet = UNet1DConditionModel(**inputs.UNET_KW).eval().cuda()
for prec in (sys.argv[1:] or ["bf16"]):
    unet.native_precision = prec
    x, cond, ehs = (t.cuda() for t in inputs.make_inputs(2, 72, 40, seed=5))
    nat = diffvits_b200.NativeDenoiser(unet, cond, ehs)
    out = nat(x, torch.tensor([500.0], device="cuda"))
    lat = diffvits_b200.sample_native(unet, x, cond, ehs, sampler="unipc", steps=3, betas=inputs.betas_linear(), native=nat, graph=False)
    torch.cuda.synchronize()
    nat.engine.check_finite()
    print(prec, "ok", float(out.abs().mean()), float(lat.abs().mean()))
    nat.close()
