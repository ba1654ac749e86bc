"""ORACLE (test infrastructure, not product code): a functional restatement of the reference's
`UNet1DConditionModel.forward` in plain torch ops on a reference-layout `state_dict`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product path (diff-vits_b200/) never does: it calls libdiffvits_b200 and raises when that is unavailable.

Parity pin: the reference ships no tests or golden vectors, so this restatement is pinned
against outputs of the reference itself, imported from /root/reference in the build container:
tests/golden/*.npz are produced by oracle/make_golden.py and tests/test_oracle.py checks the oracle against them on
every run (and against the live reference when /root/reference is present).

Each function cites the reference lines it follows (paths relative to the reference tree).

`emulate_bf16=True` rounds tensors to bfloat16 at the points where the CUDA engine stores BF16 operands
(activations written by a kernel, GEMM weights, attention probabilities) while keeping fp32 accumulation, fp32
norm statistics and fp32 time embedding. It is the tight comparison target for the CUDA path: differences that
remain are accumulation order and 1-ulp rounding flips, not operand precision.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def default_cfg() -> dict:
    """Diffusion_Encoder's UNet (model3.py:887-896 with config.json:40-46)."""
    return dict(in_channels=228, out_channels=100, block_out_channels=(128, 256, 384, 512), layers_per_block=2,
                norm_num_groups=8, norm_eps=1e-5, cross_attention_dim=128, num_heads=8)


class _Ctx:
    def __init__(self, sd: Dict[str, Tensor], cfg: dict, emulate_bf16: bool):
        self.sd, self.cfg, self.emu = sd, cfg, emulate_bf16
        self.trace: Dict[str, Tensor] = {}

    def r(self, x: Tensor) -> Tensor:
        """Round to bf16 where the engine stores a BF16 tensor."""
        return x.to(torch.bfloat16).to(torch.float32) if self.emu else x

    def w(self, key: str) -> Tensor:
        return self.sd[key].to(torch.float32)

    def wr(self, key: str) -> Tensor:
        """GEMM weight (BF16 operand in the engine)."""
        return self.r(self.w(key))


# ---------------------------------------------------------------------------------------------- embeddings
def timestep_embedding(t: Tensor, dim: int) -> Tensor:
    """embeddings.py:24-64 with flip_sin_to_cos=True, freq_shift=0 (unet_1d_condition.py:157-158, 275)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device)
    exponent = exponent / half
    arg = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def time_mlp(c: _Ctx, temb: Tensor) -> Tensor:
    """TimestepEmbedding: Linear -> SiLU -> Linear (embeddings.py:192-197)."""
    h = F.linear(temb, c.w("time_embedding.linear_1.weight"), c.w("time_embedding.linear_1.bias"))
    return F.linear(F.silu(h), c.w("time_embedding.linear_2.weight"), c.w("time_embedding.linear_2.bias"))


def text_time_embedding(c: _Ctx, ehs: Tensor) -> Tensor:
    """TextTimeEmbedding + AttentionPooling (embeddings.py:429-434, 512-546), 64 heads. fp32 (hoisted in the engine)."""
    p = "add_embedding."
    x = F.layer_norm(ehs, (ehs.shape[-1],), c.w(p + "norm1.weight"), c.w(p + "norm1.bias"), 1e-5)
    bs, length, width = x.shape
    heads = 64
    d = width // heads
    cls = x.mean(dim=1, keepdim=True) + c.w(p + "pool.positional_embedding")
    seq = torch.cat([cls, x], dim=1)
    q = F.linear(cls, c.w(p + "pool.q_proj.weight"), c.w(p + "pool.q_proj.bias"))
    k = F.linear(seq, c.w(p + "pool.k_proj.weight"), c.w(p + "pool.k_proj.bias"))
    v = F.linear(seq, c.w(p + "pool.v_proj.weight"), c.w(p + "pool.v_proj.bias"))
    scale = 1 / math.sqrt(math.sqrt(d))
    qh = (q * scale).view(bs, 1, heads, d).transpose(1, 2)
    kh = (k * scale).view(bs, length + 1, heads, d).transpose(1, 2)
    vh = v.view(bs, length + 1, heads, d).transpose(1, 2)
    wgt = torch.softmax(qh @ kh.transpose(-1, -2), dim=-1)
    pooled = (wgt @ vh).transpose(1, 2).reshape(bs, width)
    y = F.linear(pooled, c.w(p + "proj.weight"), c.w(p + "proj.bias"))
    return F.layer_norm(y, (y.shape[-1],), c.w(p + "norm2.weight"), c.w(p + "norm2.bias"), 1e-5)


# ---------------------------------------------------------------------------------------------- building blocks
def group_norm(c: _Ctx, x: Tensor, key: str, eps: float) -> Tensor:
    return F.group_norm(x, c.cfg["norm_num_groups"], c.w(key + ".weight"), c.w(key + ".bias"), eps)


def conv(c: _Ctx, x: Tensor, key: str, stride: int = 1, padding: int = 1) -> Tensor:
    """LoRACompatibleConv.forward without LoRA = F.conv1d (lora.py:99-104)."""
    return F.conv1d(x, c.wr(key + ".weight"), c.w(key + ".bias"), stride=stride, padding=padding)


def resnet(c: _Ctx, x: Tensor, emb_act: Tensor, p: str) -> Tensor:
    """ResnetBlock2D.forward, scale_shift variant (resnet.py:591-641). x is NCT."""
    eps = c.cfg["norm_eps"]
    h = c.r(F.silu(group_norm(c, x, p + ".norm1", eps)))                              # :596-600
    h = c.r(conv(c, h, p + ".conv1"))                                                 # :612
    temb = F.linear(emb_act, c.w(p + ".time_emb_proj.weight"), c.w(p + ".time_emb_proj.bias"))[:, :, None]  # :614-617
    h = group_norm(c, h, p + ".norm2", eps)                                           # :625
    scale, shift = torch.chunk(temb, 2, dim=1)                                        # :627-629
    h = c.r(F.silu(h * (1 + scale) + shift))                                          # :629-631
    h = F.conv1d(h, c.wr(p + ".conv2.weight"), c.w(p + ".conv2.bias"), padding=1)     # :634
    if (p + ".conv_shortcut.weight") in c.sd:                                         # :636-637
        x = F.conv1d(x, c.wr(p + ".conv_shortcut.weight"), c.w(p + ".conv_shortcut.bias"))
    return c.r(x + h)                                                                 # :639 (output_scale_factor 1)


def attention(c: _Ctx, x: Tensor, ctx: Optional[Tensor], bias: Optional[Tensor], p: str) -> Tensor:
    """Attention + AttnProcessor2_0 (attention_processor.py:980-1052). x [B, L, C]; ctx [B, Lk, Dc] or None;
    bias: additive [B, 1, 1, Lk] or None. Returns to_out(attn) WITHOUT the residual."""
    heads = c.cfg["num_heads"]
    src = x if ctx is None else ctx
    q = c.r(F.linear(x, c.wr(p + ".to_q.weight")))                                    # :1011
    k = c.r(F.linear(src, c.wr(p + ".to_k.weight")))                                  # :1018
    v = c.r(F.linear(src, c.wr(p + ".to_v.weight")))                                  # :1019
    B, L, C = q.shape
    d = C // heads
    qh = q.view(B, L, heads, d).transpose(1, 2)
    kh = k.view(B, -1, heads, d).transpose(1, 2)
    vh = v.view(B, -1, heads, d).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * (d ** -0.5)                                     # SDPA :1032, scale :79
    if bias is not None:
        s = s + bias
    if c.emu:
        m = s.max(dim=-1, keepdim=True).values
        pu = torch.exp(s - m)
        o = (c.r(pu) @ vh) / pu.sum(dim=-1, keepdim=True)     # engine: bf16 P operand, fp32 normaliser
    else:
        o = torch.softmax(s, dim=-1) @ vh
    o = c.r(o.transpose(1, 2).reshape(B, L, C))                                        # :1036
    return F.linear(o, c.wr(p + ".to_out.0.weight"), c.w(p + ".to_out.0.bias"))        # :1040


def transformer(c: _Ctx, y: Tensor, ehs: Tensor, bias: Optional[Tensor], p: str) -> Tensor:
    """Transformer2DModel.forward (transformer_1d.py:253-295) + BasicTransformerBlock.forward (attention.py:142-201)."""
    res = y
    g = c.r(group_norm(c, y, p + ".norm", 1e-6))                                       # transformer_1d.py:256
    z = c.r(conv(c, g, p + ".proj_in", padding=0)).permute(0, 2, 1)                    # :258-260
    tb = p + ".transformer_blocks.0"
    C = z.shape[-1]

    def ln(x, key):
        return c.r(F.layer_norm(x, (C,), c.w(key + ".weight"), c.w(key + ".bias"), 1e-5))

    z = c.r(attention(c, ln(z, tb + ".norm1"), None, None, tb + ".attn1") + z)         # attention.py:152-165
    z = c.r(attention(c, ln(z, tb + ".norm2"), c.r(ehs), bias, tb + ".attn2") + z)     # :168-177
    n3 = ln(z, tb + ".norm3")                                                          # :180
    proj = F.linear(n3, c.wr(tb + ".ff.net.0.proj.weight"), c.w(tb + ".ff.net.0.proj.bias"))
    u, gate = proj.chunk(2, dim=-1)                                                    # GEGLU attention.py:299-301
    gg = c.r(u * F.gelu(gate))
    z = c.r(F.linear(gg, c.wr(tb + ".ff.net.2.weight"), c.w(tb + ".ff.net.2.bias")) + z)   # :199-201
    zc = z.permute(0, 2, 1)                                                            # transformer_1d.py:289
    out = F.conv1d(zc, c.wr(p + ".proj_out.weight"), c.w(p + ".proj_out.bias"))        # :290
    return c.r(out + res)                                                              # :295


def upsample(c: _Ctx, x: Tensor, size: int, key: str) -> Tensor:
    """Upsample2D.forward with output_size (resnet.py:157-170): nearest resize then conv k3 p1."""
    up = F.interpolate(x, size=size, mode="nearest")
    if c.emu and size in (2 * x.shape[-1], 2 * x.shape[-1] - 1):
        # the engine folds the nearest resize into the weights (operand rounding points follow its packing):
        # to 2L rows: even rows [W0 | W1+W2 | 0], odd rows [0 | W0+W1 | W2] over (x[p-1], x[p], x[p+1]);
        # to 2L - 1 rows: even rows keep W0, W1, W2 one by one (the last one loses its W2 tap to the padding)
        w = c.w(key + ".weight")
        w0, w1, w2 = w[:, :, 0], w[:, :, 1], w[:, :, 2]
        z = torch.zeros_like(w0)
        b = c.w(key + ".bias")
        wo = c.r(torch.stack([z, w0 + w1, w2], dim=-1))
        yo = F.conv1d(x, wo, b, padding=1)
        if size % 2 == 0:
            we = c.r(torch.stack([w0, w1 + w2, z], dim=-1))
            y = torch.stack([F.conv1d(x, we, b, padding=1), yo], dim=-1).reshape(x.shape[0], -1, size)
        else:
            y = F.conv1d(up, c.r(w), b, padding=1)
            y[..., 1::2] = yo[..., :size // 2]
        return c.r(y)
    return c.r(conv(c, up, key))


# ---------------------------------------------------------------------------------------------- the network
@torch.no_grad()
def unet_forward(sd: Dict[str, Tensor], sample: Tensor, timestep: Tensor, ehs: Tensor,
                 mask: Optional[Tensor] = None, cfg: Optional[dict] = None, emulate_bf16: bool = False,
                 trace: Optional[Dict[str, Tensor]] = None, teacher: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """UNet1DConditionModel.forward (unet_1d_condition.py:743-1037).

    sample [B, Cin, T] fp32; timestep [B] (or [1]) float; ehs [B, Tp, Dc]; mask: bool [B, Tp] keep-mask or None.
    Returns [B, out_channels, T] fp32.

    trace: filled with every block output (reference module path -> [B, C, L]).
    teacher: block outputs of ANOTHER evaluation (the CUDA engine's, same names, [B, C, L]). Each block's own output still
    goes to `trace`, but everything downstream of a block consumes the teacher's tensor instead: block k is evaluated on
    the engine's input of block k, so its error does not accumulate with depth (teacher-forced per-block parity)."""
    cfg = cfg or default_cfg()
    c = _Ctx(sd, cfg, emulate_bf16)
    boc = tuple(cfg["block_out_channels"])
    n, lpb = len(boc), cfg["layers_per_block"]
    B = sample.shape[0]
    sample = sample.to(torch.float32)
    ehs = ehs.to(torch.float32)
    bias = None
    if mask is not None:                                                               # :816-818
        bias = ((1 - mask.to(torch.float32)) * -10000.0)[:, None, None, :]
    t = timestep.to(torch.float32).reshape(-1).expand(B)                               # :839
    emb = time_mlp(c, timestep_embedding(t, boc[0]))                                   # :841-848
    if "add_embedding.proj.weight" in sd:
        emb = emb + text_time_embedding(c, ehs)                                        # :869-870, 918
    emb_act = F.silu(emb)                                                              # resnet.py:614-616
    def tr(name, v):
        if trace is not None:
            trace[name] = v
        if teacher is not None:
            return teacher[name].to(v.dtype)
        return v

    h = tr("conv_in", c.r(conv(c, c.r(sample), "conv_in")))                            # :943
    skips = [h]
    for i in range(n):                                                                 # :950-976
        p = f"down_blocks.{i}"
        for j in range(lpb):
            h = tr(f"{p}.resnets.{j}", resnet(c, h, emb_act, f"{p}.resnets.{j}"))
            if i < n - 1:
                h = tr(f"{p}.attentions.{j}", transformer(c, h, ehs, bias, f"{p}.attentions.{j}"))
            skips.append(h)
        if i < n - 1:
            h = tr(f"{p}.downsamplers.0.conv", c.r(conv(c, h, f"{p}.downsamplers.0.conv", stride=2)))  # resnet.py:214-223
            skips.append(h)
    h = tr("mid_block.resnets.0", resnet(c, h, emb_act, "mid_block.resnets.0"))        # unet_1d_blocks.py:611-622
    h = tr("mid_block.attentions.0", transformer(c, h, ehs, bias, "mid_block.attentions.0"))
    h = tr("mid_block.resnets.1", resnet(c, h, emb_act, "mid_block.resnets.1"))
    for i in range(n):                                                                 # :1003-1026
        p = f"up_blocks.{i}"
        for j in range(lpb + 1):
            h = torch.cat([h, skips.pop()], dim=1)                                     # unet_1d_blocks.py:2085, 2187
            h = tr(f"{p}.resnets.{j}", resnet(c, h, emb_act, f"{p}.resnets.{j}"))
            if i > 0:
                h = tr(f"{p}.attentions.{j}", transformer(c, h, ehs, bias, f"{p}.attentions.{j}"))
        if i < n - 1:
            h = tr(f"{p}.upsamplers.0.conv",
                   upsample(c, h, skips[-1].shape[-1], f"{p}.upsamplers.0.conv"))      # :1011-1012 (size forwarded)
    h = c.r(F.silu(group_norm(c, h, "conv_norm_out", cfg["norm_eps"])))                # :1029-1031
    return conv(c, h, "conv_out")                                                      # :1032
