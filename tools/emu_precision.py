"""CPU emulation of where 16-bit rounding error comes from (test infrastructure; uses the oracle): the UNet forward with every
rounding site of the engine named, so that single site classes can be left un-rounded. Prints rel-L2 vs the FP32 oracle for
each policy at the C1 shape.   python tools/emu_precision.py [fp16|bf16]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.nn.functional as F

from oracle import inputs, unet_oracle as uo

DT = torch.float16 if (len(sys.argv) < 2 or sys.argv[1] == "fp16") else torch.bfloat16


class Ctx(uo._Ctx):
    def __init__(self, sd, cfg, exact):
        super().__init__(sd, cfg, True)
        self.exact = set(exact)

    def rs(self, x, site):
        return x if site in self.exact else x.to(DT).to(torch.float32)

    def r(self, x):                       # sites the copied functions below do not name: classified by the calling line of the oracle
        import inspect
        fr = inspect.currentframe().f_back
        src = (inspect.getframeinfo(fr).code_context or [""])[0]
        if "c.r(sample)" in src:
            return self.rs(self.rs(x, "conv_in_out"), "noop") if False else self._conv_in(x, src)
        if "conv_norm_out" in src:
            return self.rs(x, "final_norm")
        if "downsamplers" in src:
            return self.rs(x, "ds_out")
        if fr.f_code.co_name == "upsample":
            return self.rs(x, "us_out")
        return self.rs(x, "misc_other")

    def _conv_in(self, x, src):
        # the line is  c.r(conv(c, c.r(sample), "conv_in")) : the inner call rounds the input sample, the outer the conv_in output
        self._n_conv_in = getattr(self, "_n_conv_in", 0) + 1
        return self.rs(x, "in_sample" if self._n_conv_in == 1 else "conv_in_out")

    def wr(self, key):
        return self.rs(self.w(key), "w")


def resnet(c, x, emb_act, p):
    eps = c.cfg["norm_eps"]
    h = c.rs(F.silu(uo.group_norm(c, x, p + ".norm1", eps)), "norm_out")
    h = c.rs(uo.conv(c, h, p + ".conv1"), "conv1_out")
    temb = F.linear(emb_act, c.w(p + ".time_emb_proj.weight"), c.w(p + ".time_emb_proj.bias"))[:, :, None]
    h = uo.group_norm(c, h, p + ".norm2", eps)
    scale, shift = torch.chunk(temb, 2, dim=1)
    h = c.rs(F.silu(h * (1 + scale) + shift), "norm_out")
    h = F.conv1d(h, c.wr(p + ".conv2.weight"), c.w(p + ".conv2.bias"), padding=1)
    if (p + ".conv_shortcut.weight") in c.sd:
        x = F.conv1d(x, c.wr(p + ".conv_shortcut.weight"), c.w(p + ".conv_shortcut.bias"))
    return c.rs(x + h, "block_out")


def attention(c, x, ctx, bias, p):
    heads = c.cfg["num_heads"]
    src = x if ctx is None else ctx
    q = c.rs(F.linear(x, c.wr(p + ".to_q.weight")), "qkv")
    k = c.rs(F.linear(src, c.wr(p + ".to_k.weight")), "qkv")
    v = c.rs(F.linear(src, c.wr(p + ".to_v.weight")), "qkv")
    B, L, C = q.shape
    d = C // heads
    qh, kh, vh = (t.view(B, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    s = (qh @ kh.transpose(-1, -2)) * (d ** -0.5)
    if bias is not None:
        s = s + bias
    m = s.max(dim=-1, keepdim=True).values
    pu = torch.exp(s - m)
    o = (c.rs(pu, "attn_p") @ vh) / pu.sum(dim=-1, keepdim=True)
    o = c.rs(o.transpose(1, 2).reshape(B, L, C), "attn_o")
    return F.linear(o, c.wr(p + ".to_out.0.weight"), c.w(p + ".to_out.0.bias"))


def transformer(c, y, ehs, bias, p):
    res = y
    g = c.rs(uo.group_norm(c, y, p + ".norm", 1e-6), "norm_out")
    z = c.rs(uo.conv(c, g, p + ".proj_in", padding=0), "z").permute(0, 2, 1)
    tb = p + ".transformer_blocks.0"
    C = z.shape[-1]

    def ln(x, key):
        return c.rs(F.layer_norm(x, (C,), c.w(key + ".weight"), c.w(key + ".bias"), 1e-5), "norm_out")

    z = c.rs(attention(c, ln(z, tb + ".norm1"), None, None, tb + ".attn1") + z, "z")
    z = c.rs(attention(c, ln(z, tb + ".norm2"), c.rs(ehs, "ehs"), bias, tb + ".attn2") + z, "z")
    n3 = ln(z, tb + ".norm3")
    proj = F.linear(n3, c.wr(tb + ".ff.net.0.proj.weight"), c.w(tb + ".ff.net.0.proj.bias"))
    u, gate = proj.chunk(2, dim=-1)
    gg = c.rs(u * F.gelu(gate), "gg")
    z = c.rs(F.linear(gg, c.wr(tb + ".ff.net.2.weight"), c.w(tb + ".ff.net.2.bias")) + z, "z")
    out = F.conv1d(z.permute(0, 2, 1), c.wr(p + ".proj_out.weight"), c.w(p + ".proj_out.bias"))
    return c.rs(out + res, "block_out")


def run(sd, sample, t, ehs, exact):
    orig = (uo.resnet, uo.transformer, uo._Ctx)
    uo.resnet, uo.transformer = resnet, transformer
    uo._Ctx = lambda sd_, cfg_, emu_: Ctx(sd_, cfg_, exact)
    try:
        return uo.unet_forward(sd, sample, t, ehs, emulate_bf16=True)
    finally:
        uo.resnet, uo.transformer, uo._Ctx = orig


def main():
    import diffvits_b200
    from diffvits_b200.synthetic import UNET_KW
    diffvits_b200.install_dropin()
    from unet1d.unet_1d_condition import UNet1DConditionModel
    torch.manual_seed(1234)
    unet = UNet1DConditionModel(**UNET_KW).eval()
    sd = {k: v.detach() for k, v in unet.state_dict().items()}
    x_T, cond, ehs = inputs.make_inputs(1, 200, 282)
    sample = torch.cat([x_T, cond], 1)
    t = torch.tensor([500.0])
    torch.set_num_threads(os.cpu_count())
    ref = uo.unet_forward(sd, sample, t, ehs)
    sites = ["w", "norm_out", "conv1_out", "block_out", "z", "qkv", "attn_p", "attn_o", "gg", "in_sample", "conv_in_out", "ds_out",
             "us_out", "final_norm", "ehs", "misc_other"]
    policies = [("all sites rounded", [])] + [(f"exact: {s}", [s]) for s in sites] + [
        ("exact: block_out + z (= fp16s)", ["block_out", "z"]),
        ("exact: block_out + z + conv1_out", ["block_out", "z", "conv1_out"]),
        ("exact: w + block_out + z", ["w", "block_out", "z"]),
        ("exact: every activation (weights only rounded)", [s for s in sites if s != "w"]),
        ("exact: in_sample + final_norm + ehs", ["in_sample", "final_norm", "ehs"]),
        ("exact: in_sample + final_norm + ehs + conv_in_out + ds_out + us_out", ["in_sample", "final_norm", "ehs", "conv_in_out", "ds_out", "us_out"]),
    ]
    print(f"operand format {DT}")
    for name, exact in policies:
        out = run(sd, sample, t, ehs, exact)
        rel = float((out - ref).norm() / ref.norm())
        print(f"{name:50s} rel-L2 {rel:.3e}")


if __name__ == "__main__":
    main()
