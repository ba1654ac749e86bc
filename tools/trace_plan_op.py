# Needs the debug library: DV_DEBUG=1 python -m diffvits_b200.build ; run with DV_DEBUG_LIB=1
"""CTA-0 timeline of chosen tap-GEMM ops INSIDE a real denoiser call at the C2 shapes (dv_debug_trace_op): shows the fused
output-GroupNorm tail (pass 1 done -> item complete -> statistics folded -> tile done).   python tools/trace_plan_op.py [substr ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import diffvits_b200
from diffvits_b200 import _lib, synthetic as inputs
from trace_gemm import NAMES, REGION  # noqa: E402  (event names; importing runs nothing: guarded below)


def main():
    pats = sys.argv[1:] or ["down_blocks.1.resnets.0.conv1", "down_blocks.1.resnets.0.conv2", "down_blocks.1.attentions.0.proj_out",
                            "down_blocks.0.resnets.0.conv1", "down_blocks.1.attentions.0.transformer_blocks.0.attn1.qkv"]
    lib = _lib.load()
    diffvits_b200.install_dropin()
    from unet1d.unet_1d_condition import UNet1DConditionModel
    torch.manual_seed(1234)
    unet = UNet1DConditionModel(**inputs.UNET_KW).eval().cuda()
    x, cond, ehs = (t.cuda() for t in inputs.make_inputs(32, 1000, 282))
    nat = diffvits_b200.NativeDenoiser(unet, cond, ehs)
    eng = nat.engine.engines[0]
    ops = eng.op_table()
    t = torch.tensor([500.0], device="cuda")
    for _ in range(3):
        nat(x, t)
    torch.cuda.synchronize()
    for pat in pats:
        idx = [i for i, o in enumerate(ops) if o["tag"].split(" ")[0].endswith(pat) and o["kind"] == 0]
        if not idx:
            print("no op matches", pat)
            continue
        i = idx[0]
        buf = torch.zeros(1 + 4 * REGION, dtype=torch.int64, device="cuda")
        assert lib.dv_debug_trace_op(eng._plan, i, buf.data_ptr()) == 0
        nat(x, t)
        torch.cuda.synchronize()
        buf.zero_()
        nat(x, t)
        torch.cuda.synchronize()
        lib.dv_debug_trace_op(eng._plan, i, None)
        h = buf.cpu().tolist()
        recs = []
        for role in range(3):
            base = 1 + role * REGION
            n = h[base] & ((1 << 64) - 1)
            for v in h[base + 1: base + 1 + n]:
                v &= (1 << 64) - 1
                recs.append((v & 0xFFFFFFFFFF, role, (v >> 44) & 0xFFFF, (v >> 40) & 0xF))
        recs.sort()
        print(f"== op {i}: {ops[i]['tag']}")
        t0 = recs[0][0]
        for tt, role, tile, ev in recs:
            print(f"     {tt - t0:7d} clk  {NAMES.get((role, ev), (role, ev))}" + (f"  ({tile})" if ev < 8 and tile else ""))


if __name__ == "__main__":
    main()
