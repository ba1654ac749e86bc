"""Per-kernel parity on the B200 (through the C ABI) against plain fp32 torch math on the same BF16-rounded inputs.
Tolerances: a BF16 result is one rounding (2^-9 relative) away from the fp32 value, so rel-L2 <= 4e-3 and max-abs
<= 2^-7 * max|ref|; fp32 kernels (sampler updates) <= 1e-6."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gpu_util
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return gpu_util


def _check_bf16(U, got, ref, rel=4e-3):
    assert torch.isfinite(got.float()).all()
    assert U.rel_l2(got.float(), ref) < rel, U.rel_l2(got.float(), ref)
    assert U.max_abs(got.float(), ref) <= 2 ** -7 * float(ref.abs().max()) + 1e-3


def _rand(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("B,L,K,N", [(2, 300, 128, 128), (1, 125, 512, 256), (3, 1000, 64, 384), (1, 25, 1024, 512)])
def test_gemm_linear(U, B, L, K, N):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = _rand((B, L, K), g)
    w = _rand((N, K), g, K ** -0.5)
    bias = torch.randn(N, generator=g, device="cuda")
    res = _rand((B, L, N), g)
    ref = a.float() @ w.float().t() + bias + res.float()
    got = U.gemm([a], [(0, 0, 0, K)], w, N, bias=bias, res=res)
    _check_bf16(U, got, ref)
    raw = U.gemm([a], [(0, 0, 0, K)], w, N, bias=bias, impl=1)      # CUDA-core reference kernel, raw fp32
    assert U.rel_l2(raw[:, :, :N], a.float() @ w.float().t() + bias) < 1e-5


@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_tile_width_and_groupnorm_partials(U, bn):
    """Both N-tile instantiations (persistent schedule over many tiles) + the epilogue's GroupNorm partial sums:
    per (32-row chunk, 16-channel block) sum / sum of squares of the fp32 accumulator values (the stored bf16 values
    differ from them by rounding noise only, hence the 1e-3 gate against sums of the stored tensor)."""
    g = torch.Generator(device="cuda").manual_seed(21)
    B, L, K, N = 5, 1000, 384, 512            # 5 x 8 row tiles x 2-4 column tiles: several tiles per CTA... on 148 SMs
    a = _rand((B, L, K), g)
    w = _rand((N, K), g, K ** -0.5)
    bias = torch.randn(N, generator=g, device="cuda")
    res = _rand((B, L, N), g)
    ref = a.float() @ w.float().t() + bias + res.float()
    got, st = U.gemm([a], [(0, 0, 0, K)], w, N, bias=bias, res=res, bn=bn, stats=True)
    _check_bf16(U, got, ref)
    rows = (L + 127) // 128 * 128
    y = torch.zeros(B, rows, N, device="cuda")
    y[:, :L] = got.float()
    blk = y.reshape(B, rows // 32, 32, N // 16, 16)
    want = torch.stack([blk.sum(dim=(2, 4)), (blk * blk).sum(dim=(2, 4))], dim=-1)
    assert torch.isfinite(st).all()
    assert U.rel_l2(st, want) < 1e-3
    # many more tiles than SMs: every CTA loops (B=40 -> 320 row tiles x N/bn)
    B2 = 40
    a2 = _rand((B2, L, 128), g)
    w2 = _rand((256, 128), g, 128 ** -0.5)
    got2 = U.gemm([a2], [(0, 0, 0, 128)], w2, 256, bn=bn)
    _check_bf16(U, got2, a2.float() @ w2.float().t())


def test_gemm_tile_width_192(U):
    """N = 384 (level 2) as two 192-wide tiles: one wave of 128 x 192 tiles instead of 1.3 waves of 128 x 128.
    conv k3 with residual and GroupNorm partials, and a plain linear, against fp32 math and the 128-wide kernel."""
    g = torch.Generator(device="cuda").manual_seed(22)
    B, L, C, N = 3, 250, 384, 384
    x = _rand((B, L, C), g)
    w3 = _rand((N, C, 3), g, (3 * C) ** -0.5)
    bias = torch.randn(N, generator=g, device="cuda")
    res = _rand((B, L, N), g)
    ref = F.conv1d(x.float().transpose(1, 2), w3.float(), bias, padding=1).transpose(1, 2) + res.float()
    wp = w3.permute(0, 2, 1).reshape(N, 3 * C).contiguous()
    segs = [(0, -1, 0, C), (0, 0, 0, C), (0, 1, 0, C)]
    got, st = U.gemm([x], segs, wp, N, bias=bias, res=res, bn=192, stats=True)
    _check_bf16(U, got, ref)
    got128, st128 = U.gemm([x], segs, wp, N, bias=bias, res=res, bn=128, stats=True)
    assert torch.equal(got, got128)                      # same K order, same accumulation: bit-identical
    assert torch.equal(st, st128)
    a = _rand((B, L, 4 * C), g)
    w = _rand((N, 4 * C), g, (4 * C) ** -0.5)
    got = U.gemm([a], [(0, 0, 0, 4 * C)], w, N, bn=192)
    _check_bf16(U, got, a.float() @ w.float().t())
    # automatic choice at this shape must agree too (whichever tile it picks)
    assert torch.equal(got, U.gemm([a], [(0, 0, 0, 4 * C)], w, N))


@pytest.mark.parametrize("B,L,Cin,Cout", [(2, 257, 128, 128), (1, 200, 256, 384), (2, 1000, 384, 128), (2, 26, 512, 512)])
def test_gemm_conv3(U, B, L, Cin, Cout):
    """conv1d k3 s1 p1: three row-shifted K segments, zero fill outside [0, L)."""
    g = torch.Generator(device="cuda").manual_seed(2)
    x = _rand((B, L, Cin), g)
    w = (torch.randn(Cout, Cin, 3, generator=g, device="cuda") * (3 * Cin) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g, device="cuda")
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=1).transpose(1, 2)
    wp = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous()
    got = U.gemm([x], [(0, -1, 0, Cin), (0, 0, 0, Cin), (0, 1, 0, Cin)], wp, Cout, bias=bias)
    _check_bf16(U, got, ref)


@pytest.mark.parametrize("B,L,Cin,Cout,bn", [(2, 257, 128, 128, 0), (1, 200, 256, 384, 192), (2, 1000, 384, 128, 64),
                                             (2, 26, 512, 512, 256), (40, 1000, 128, 128, 0), (1, 1, 64, 128, 0)])
def test_gemm_conv3_shared_taps(U, B, L, Cin, Cout, bn):
    """conv1d k3 s1 p1 with ONE staged 136-row box per channel panel and three row-shifted tensor-core reads
    (profiles/r2b_umma_row_shift_probe.txt) instead of three loads: same contract as three single-tap segments."""
    g = torch.Generator(device="cuda").manual_seed(2)
    x = _rand((B, L, Cin), g)
    w = (torch.randn(Cout, Cin, 3, generator=g, device="cuda") * (3 * Cin) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g, device="cuda")
    res = _rand((B, L, Cout), g)
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=1).transpose(1, 2) + res.float()
    wp = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous()
    got, st = U.gemm([x], [(0, -1, 0, Cin, 3)], wp, Cout, bias=bias, res=res, bn=bn, stats=True)
    _check_bf16(U, got, ref)
    three, st3 = U.gemm([x], [(0, -1, 0, Cin), (0, 0, 0, Cin), (0, 1, 0, Cin)], wp, Cout, bias=bias, res=res, bn=bn, stats=True)
    assert U.rel_l2(got.float(), three.float()) < 2e-3          # same products, different accumulation order
    assert U.rel_l2(st, st3) < 1e-3


@pytest.mark.parametrize("B,L,C0,C1,N,silu,ss,bn", [(3, 250, 256, 128, 256, True, False, 0), (2, 1000, 128, 0, 128, True, True, 0),
                                                     (33, 500, 256, 0, 256, True, True, 256), (2, 125, 512, 384, 384, True, False, 192),
                                                     (2, 26, 512, 512, 512, True, True, 0), (1, 7, 128, 0, 128, False, False, 64)])
def test_gemm_conv3_with_fused_groupnorm(U, B, L, C0, C1, N, silu, ss, bn):
    """GroupNorm (+ time scale/shift, + SiLU) of the conv's input applied in shared memory between the TMA load and the
    MMA (resnet.py:596-600, 625-631 then conv1 / conv2): the normalised tensor is never written. Two sources = the
    never-materialised torch.cat([h, skip]) with groups that straddle the seam. Reference: torch GroupNorm on the
    concatenated tensor, operand rounded to bf16 as the kernel stores it, fp32 conv."""
    g = torch.Generator(device="cuda").manual_seed(7)
    Cc, G = C0 + C1, 8
    x0 = _rand((B, L, C0), g, 1.7) + 0.3
    x1 = (_rand((B, L, C1), g, 0.6) - 0.2) if C1 else None
    gamma = torch.randn(Cc, generator=g, device="cuda") * 0.3 + 1.0
    beta = torch.randn(Cc, generator=g, device="cuda") * 0.2
    ssv = (torch.randn(B, 2 * Cc, generator=g, device="cuda") * 0.3) if ss else None
    w = (torch.randn(N, Cc, 3, generator=g, device="cuda") * (3 * Cc) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda")
    xc = x0.float() if x1 is None else torch.cat([x0, x1], 2).float()
    y = F.group_norm(xc.transpose(1, 2), G, gamma, beta, 1e-5)
    if ss:
        y = y * (1 + ssv[:, :Cc, None]) + ssv[:, Cc:, None]
    if silu:
        y = F.silu(y)
    y = y.to(torch.bfloat16).float()
    ref = F.conv1d(y, w.float(), bias, padding=1).transpose(1, 2)
    wp = w.permute(0, 2, 1).reshape(N, 3 * Cc).contiguous()
    np_ = Cc // 64
    srcs = [x0] + ([x1] if C1 else [])
    segs = [(0, -1, 0, C0, 3, 0, np_, 1, 0)] + ([(1, -1, 0, C1, 3, C0 // 64, np_, 1, C0)] if C1 else [])
    norm = dict(stats=[U.gn_partials(t) for t in srcs], cover=[t.shape[2] for t in srcs], C=Cc, L=L, G=G, eps=1e-5,
                silu=silu, gamma=gamma, beta=beta, scale_shift=ssv)
    got = U.gemm(srcs, segs, wp, N, bias=bias, bn=bn, norm=norm)
    _check_bf16(U, got, ref)
    # the stand-alone GroupNorm kernel + plain conv give the same operand: results agree to accumulation order
    a = U.groupnorm(x0, x1, G, gamma, beta, 1e-5, silu, ssv)
    unf = U.gemm([a], [(0, -1, 0, Cc, 3)], wp, N, bias=bias, bn=bn)
    assert U.rel_l2(got.float(), unf.float()) < 2e-3


@pytest.mark.parametrize("B", [2, 40])
def test_gemm_linear_with_fused_groupnorm_and_raw_shortcut(U, B):
    """1x1 GEMM whose operand is GroupNorm(x) without SiLU (transformer_1d.py:256-258 norm + proj_in), and a conv whose
    normalised operand shares the launch with RAW shortcut sources (resnet conv2 + conv_shortcut over (h | skip)).
    B = 40: several tiles per persistent CTA, more raw panels per tile than A slots (the normalised and the raw panels
    complete different barriers of a slot; a transform warp that lost step with the ring would trap here)."""
    g = torch.Generator(device="cuda").manual_seed(8)
    L, Cc, N, G = 500, 256, 256, 8
    x = _rand((B, L, Cc), g, 2.0)
    gamma, beta = torch.randn(Cc, generator=g, device="cuda") * 0.2 + 1, torch.randn(Cc, generator=g, device="cuda") * 0.1
    w = _rand((N, Cc), g, Cc ** -0.5)
    bias = torch.randn(N, generator=g, device="cuda")
    y = F.group_norm(x.float().transpose(1, 2), G, gamma, beta, 1e-6).to(torch.bfloat16).float().transpose(1, 2)
    norm = dict(stats=[U.gn_partials(x)], cover=[Cc], C=Cc, L=L, G=G, eps=1e-6, silu=False, gamma=gamma, beta=beta)
    got = U.gemm([x], [(0, 0, 0, Cc, 1, -1, 0, 1, 0)], w, N, bias=bias, norm=norm)
    _check_bf16(U, got, y @ w.float().t() + bias)
    # conv2(SiLU(GN(u))) + shortcut(h | s): segments u (normalised, shared taps), h, s (raw)
    C1, C2 = 256, 128
    u, h, s2 = _rand((B, L, Cc), g, 1.5), _rand((B, L, C1), g), _rand((B, L, C2), g)
    w2 = (torch.randn(N, Cc, 3, generator=g, device="cuda") * (3 * Cc) ** -0.5).to(torch.bfloat16)
    wsc = (torch.randn(N, C1 + C2, 1, generator=g, device="cuda") * (C1 + C2) ** -0.5).to(torch.bfloat16)
    v = F.silu(F.group_norm(u.float().transpose(1, 2), G, gamma, beta, 1e-5)).to(torch.bfloat16).float()
    ref = (F.conv1d(v, w2.float(), None, padding=1)
           + F.conv1d(torch.cat([h, s2], 2).float().transpose(1, 2), wsc.float(), bias)).transpose(1, 2)
    wp = torch.cat([w2.permute(0, 2, 1).reshape(N, 3 * Cc), wsc[:, :, 0]], dim=1).contiguous()
    segs = [(0, -1, 0, Cc, 3, 0, Cc // 64, 1, 0), (1, 0, 0, C1), (2, 0, 0, C2)]
    norm = dict(stats=[U.gn_partials(u)], cover=[Cc], C=Cc, L=L, G=G, eps=1e-5, silu=True, gamma=gamma, beta=beta)
    got = U.gemm([u, h, s2], segs, wp, N, bias=bias, norm=norm)
    _check_bf16(U, got, ref)


def test_gemm_conv3_with_fused_shortcut_two_sources(U):
    """conv2 + 1x1 shortcut over torch.cat([h, skip]) as extra K panels from two tensors (resnet.py:634-639)."""
    g = torch.Generator(device="cuda").manual_seed(3)
    B, L, C, C1, C2 = 2, 250, 256, 256, 128
    v, h, s = _rand((B, L, C), g), _rand((B, L, C1), g), _rand((B, L, C2), g)
    w2 = (torch.randn(C, C, 3, generator=g, device="cuda") * (3 * C) ** -0.5).to(torch.bfloat16)
    wsc = (torch.randn(C, C1 + C2, 1, generator=g, device="cuda") * (C1 + C2) ** -0.5).to(torch.bfloat16)
    b = torch.randn(C, generator=g, device="cuda")
    ref = (F.conv1d(v.float().transpose(1, 2), w2.float(), None, padding=1)
           + F.conv1d(torch.cat([h, s], 2).float().transpose(1, 2), wsc.float(), b)).transpose(1, 2)
    wp = torch.cat([w2.permute(0, 2, 1).reshape(C, 3 * C), wsc[:, :, 0]], dim=1).contiguous()
    segs = [(0, -1, 0, C), (0, 0, 0, C), (0, 1, 0, C), (1, 0, 0, C1), (2, 0, 0, C2)]
    got = U.gemm([v, h, s], segs, wp, C, bias=b)
    _check_bf16(U, got, ref)


@pytest.mark.parametrize("L", [500, 251, 77])
def test_gemm_conv3_stride2_view(U, L):
    """Downsample2D conv k3 s2 p1 (resnet.py:214-223), L -> ceil(L / 2), through the pair view [L/2, 2C]: the first-half
    tap reads a view with ceil(L/2) rows, the second-half taps one with floor(L/2) rows, so for odd L the missing x[L]
    is TMA zero fill and nothing of the next batch item is read."""
    g = torch.Generator(device="cuda").manual_seed(4)
    B, C = 3, 128
    Lo = (L + 1) // 2
    x = _rand((B, L, C), g)
    w = (torch.randn(C, C, 3, generator=g, device="cuda") * (3 * C) ** -0.5).to(torch.bfloat16)
    b = torch.randn(C, generator=g, device="cuda")
    ref = F.conv1d(x.float().transpose(1, 2), w.float(), b, stride=2, padding=1).transpose(1, 2)
    assert ref.shape[1] == Lo
    wp = w.permute(0, 2, 1).reshape(C, 3 * C).contiguous()
    got = U.gemm([x, x], [(1, -1, C, C), (0, 0, 0, C), (1, 0, C, C)], wp, C, bias=b, l_out=Lo,
                 views=[(Lo, 2 * C), (L // 2, 2 * C)])
    _check_bf16(U, got, ref)


@pytest.mark.parametrize("odd", [0, 1])
def test_gemm_upsample_folded(U, odd):
    """nearest resize to 2L or 2L - 1 rows + conv k3 p1 folded into the weights over the low-res rows (resnet.py:138-173):
    [2C, 3C] over (x, -1) (x, 0) (x, +1) for 2L; [2C, 4C] over (x, -1) (x, 0) (x', 0) (x, +1) for 2L - 1, x' = x behind a
    map one row shorter, and the odd half of the last GEMM row is neither stored nor counted in the GroupNorm partials."""
    from diffvits_b200.engine import pack_slot
    g = torch.Generator(device="cuda").manual_seed(5)
    B, L, C = 3, 125, 256
    Ls = 2 * L - odd
    x = _rand((B, L, C), g)
    w = (torch.randn(C, C, 3, generator=g, device="cuda") * (3 * C) ** -0.5).to(torch.bfloat16).float()
    b = torch.randn(C, generator=g, device="cuda")
    up = F.interpolate(x.float().transpose(1, 2), size=Ls, mode="nearest")
    ref = F.conv1d(up, w, b, padding=1).transpose(1, 2)                       # [B, Ls, C]
    if odd:
        slot = dict(recipe="conv3_up2_odd", shape=[2 * C, 4 * C], keys=["w"], dtype="bf16", name="t")
        wp = pack_slot(slot, {"w": w.cpu()}).reshape(2 * C, 4 * C).cuda()
        got, st = U.gemm([x, x], [(0, -1, 0, C), (0, 0, 0, C), (1, 0, 0, C), (0, 1, 0, C)], wp, 2 * C,
                         bias=torch.cat([b, b]), views=[(L, C), (L - 1, C)], stats=True, layout=(Ls * C, C))
    else:
        slot = dict(recipe="conv3_up2", shape=[2 * C, 3 * C], keys=["w"], dtype="bf16", name="t")
        wp = pack_slot(slot, {"w": w.cpu()}).reshape(2 * C, 3 * C).cuda()
        got, st = U.gemm([x], [(0, -1, 0, C), (0, 0, 0, C), (0, 1, 0, C)], wp, 2 * C, bias=torch.cat([b, b]), stats=True,
                         layout=(Ls * C, 0))
    got = got.reshape(B, Ls, C)                                              # dense [B, Ls, C]: nothing spilled over
    _check_bf16(U, got, ref, rel=6e-3)                                        # W0+W1 is rounded once more
    # GroupNorm partials cover exactly the Ls x C stored values (folded: block kb and kb + C/16 are the same channels)
    assert torch.isfinite(st).all()
    tot = st.sum(dim=1)                                                       # [B, 2C/16, 2]
    tot = tot[:, :C // 16] + tot[:, C // 16:]
    y = got.float().reshape(B, Ls, C // 16, 16)
    want = torch.stack([y.sum(dim=(1, 3)), (y * y).sum(dim=(1, 3))], dim=-1)
    assert U.rel_l2(tot, want) < 2e-3


def test_gemm_geglu_and_ncl(U):
    from diffvits_b200.engine import pack_slot
    g = torch.Generator(device="cuda").manual_seed(6)
    B, L, C = 2, 300, 128
    a = _rand((B, L, C), g)
    w = (torch.randn(8 * C, C, generator=g, device="cuda") * C ** -0.5).to(torch.bfloat16).float()
    b = torch.randn(8 * C, generator=g, device="cuda")
    proj = a.float() @ w.t() + b
    u, gate = proj.chunk(2, dim=-1)
    ref = u * F.gelu(gate)
    wp = pack_slot(dict(recipe="geglu", shape=[8 * C, C], keys=["w"], dtype="bf16", name="t"), {"w": w.cpu()})
    bp = pack_slot(dict(recipe="geglu_vec", shape=[8 * C], keys=["b"], dtype="f32", name="t"), {"b": b.cpu()})
    got = U.gemm([a], [(0, 0, 0, C)], wp.reshape(8 * C, C).cuda(), 8 * C, bias=bp.cuda(), out_mode=2)
    _check_bf16(U, got, ref)
    # conv_out style: N=100 padded to 128, fp32 [B, N, L] store
    w3 = (torch.randn(100, C, 3, generator=g, device="cuda") * (3 * C) ** -0.5).to(torch.bfloat16)
    b3 = torch.randn(100, generator=g, device="cuda")
    ref3 = F.conv1d(a.float().transpose(1, 2), w3.float(), b3, padding=1)
    wp3 = torch.zeros(128, 3 * C, dtype=torch.bfloat16, device="cuda")
    wp3[:100] = w3.permute(0, 2, 1).reshape(100, 3 * C)
    bp3 = torch.zeros(128, device="cuda")
    bp3[:100] = b3
    got3 = U.gemm([a], [(0, -1, 0, C), (0, 0, 0, C), (0, 1, 0, C)], wp3, 100, bias=bp3, out_mode=1)
    assert U.rel_l2(got3, ref3) < 1e-5


@pytest.fixture(params=["tc", "legacy", "poly0", "poly3", "poly5"])
def attn_kernel(request, monkeypatch):
    """Every attention kernel (tcgen05 / first mma.sync / second mma.sync with 0, 3 or 5 of 8 accumulator columns on
    the FMA-pipe exp2) is checked at every head dim, whichever the dispatcher prefers."""
    if request.param.startswith("poly"):
        monkeypatch.setenv("DV_ATTN", "v2")
        monkeypatch.setenv("DV_ATTN_POLY", request.param[4:])
    else:
        monkeypatch.setenv("DV_ATTN", request.param)
    return request.param


@pytest.mark.parametrize("gain", [3.0, 12.0, 40.0])
def test_attention_wide_score_range(U, attn_kernel, gain):
    """Peaky softmax: scaled scores spread over tens to thousands of log2 units, so most exponents sit far below the
    row maximum (the polynomial exp2 path clamps at 2^-126 where MUFU.EX2 flushes to zero)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    B, H, D, L = 2, 8, 16, 300
    C = H * D
    q = (_rand((B, L, C), g).float() * gain).to(torch.bfloat16)
    kv = _rand((B, L, 2 * C), g)
    k, v = kv[:, :, :C], kv[:, :, C:]
    ref = F.scaled_dot_product_attention(*(t.float().reshape(B, L, H, D).transpose(1, 2) for t in (q, k, v)))
    ref = ref.transpose(1, 2).reshape(B, L, C)
    got = U.attention(q, k, v, H)
    assert torch.isfinite(got.float()).all()
    _check_bf16(U, got, ref, rel=8e-3)


@pytest.mark.parametrize("D,Lq,Lk", [(16, 200, 200), (32, 500, 500), (48, 250, 250), (64, 125, 125), (16, 1000, 1000),
                                     (32, 77, 40), (16, 130, 65)])
def test_attention_self(U, attn_kernel, D, Lq, Lk):
    g = torch.Generator(device="cuda").manual_seed(7)
    B, H = 2, 8
    C = H * D
    if Lq == Lk:
        qkv = _rand((B, Lq, 3 * C), g)
        q, k, v = qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:]
    else:                                   # ragged tails of the query / key tiles
        q = _rand((B, Lq, C), g)
        kv = _rand((B, Lk, 2 * C), g)
        k, v = kv[:, :, :C], kv[:, :, C:]
    ref = F.scaled_dot_product_attention(*(t.float().reshape(B, t.shape[1], H, D).transpose(1, 2) for t in (q, k, v)))
    ref = ref.transpose(1, 2).reshape(B, Lq, C)
    got = U.attention(q, k, v, H)
    _check_bf16(U, got, ref, rel=8e-3)        # + bf16 rounding of P


@pytest.mark.parametrize("D,Lq", [(16, 1000), (32, 500), (48, 250), (64, 125)])
def test_attention_cross_masked(U, attn_kernel, D, Lq):
    g = torch.Generator(device="cuda").manual_seed(8)
    B, H, Lk = 3, 8, 282
    C = H * D
    q = _rand((B, Lq, C), g)
    kv = _rand((B, Lk, 2 * C), g)
    k, v = kv[:, :, :C], kv[:, :, C:]
    lens = torch.tensor([282, 182, 1], dtype=torch.int32, device="cuda")
    bias = ((torch.arange(Lk, device="cuda")[None, :] >= lens[:, None]).float() * -10000.0)[:, None, None, :]
    ref = F.scaled_dot_product_attention(q.float().reshape(B, Lq, H, D).transpose(1, 2),
                                         k.float().reshape(B, Lk, H, D).transpose(1, 2),
                                         v.float().reshape(B, Lk, H, D).transpose(1, 2), attn_mask=bias)
    ref = ref.transpose(1, 2).reshape(B, Lq, C)
    got = U.attention(q, k, v, H, kv_len=lens)
    _check_bf16(U, got, ref, rel=8e-3)


@pytest.mark.parametrize("C0,C1,L", [(128, 0, 1000), (512, 384, 125), (384, 0, 250), (256, 128, 77)])
def test_groupnorm(U, C0, C1, L):
    g = torch.Generator(device="cuda").manual_seed(9)
    B, G = 3, 8
    x0 = _rand((B, L, C0), g, 2.0) + 0.5
    x1 = (_rand((B, L, C1), g) - 0.25) if C1 else None
    C = C0 + C1
    gamma = torch.randn(C, generator=g, device="cuda")
    beta = torch.randn(C, generator=g, device="cuda")
    ss = torch.randn(B, 2 * C + 64, generator=g, device="cuda") * 0.3
    x = torch.cat([x0, x1], 2) if C1 else x0
    gn = F.group_norm(x.float().transpose(1, 2), G, gamma, beta, 1e-5)
    ref = F.silu(gn * (1 + ss[:, :C, None]) + ss[:, C:2 * C, None]).transpose(1, 2)
    got = U.groupnorm(x0, x1, G, gamma, beta, 1e-5, True, ss)
    _check_bf16(U, got, ref)
    ref2 = F.group_norm(x.float().transpose(1, 2), G, gamma, beta, 1e-6).transpose(1, 2)
    _check_bf16(U, U.groupnorm(x0, x1, G, gamma, beta, 1e-6, False), ref2)


@pytest.mark.parametrize("C", [128, 256, 384, 512])
def test_layernorm(U, C):
    g = torch.Generator(device="cuda").manual_seed(10)
    x = _rand((2, 333, C), g, 1.5) + 0.3
    gamma = torch.randn(C, generator=g, device="cuda")
    beta = torch.randn(C, generator=g, device="cuda")
    ref = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    _check_bf16(U, U.layernorm(x, gamma, beta), ref)


def test_sampler_update_kernels(U):
    import ctypes as C
    from diffvits_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(11)
    n = 3 * 100 * 203                      # not a multiple of 4: exercises the scalar tail
    x, m1, m0, mt = (torch.randn(n, generator=g, device="cuda") for _ in range(4))
    a, b, c = 0.93, -0.21, 0.07
    out = torch.empty_like(x)
    _lib.check(lib.dv_dpmpp_update(U.vp(out), U.vp(x), U.vp(m1), U.vp(m0), a, b, c, 0, 1.0, 1.0, n, U.stream()))
    assert U.rel_l2(out, a * x + b * m1 + c * m0) < 1e-6
    _lib.check(lib.dv_dpmpp_update(U.vp(out), U.vp(x), U.vp(m1), None, a, b, c, 0, 1.0, 1.0, n, U.stream()))
    assert U.rel_l2(out, a * x + b * m1) < 1e-6
    # round trip: m1 <- (x - sig*((x - alp*m1)/sig))/alp, bit-exact vs torch fp32 evaluated in the same order
    alp, sig = 0.0063, 0.99998
    m1c = m1.clone()
    ta, tsg = torch.tensor(alp, device="cuda"), torch.tensor(sig, device="cuda")
    rt = (x - tsg * ((x - ta * m1) / tsg)) / ta
    _lib.check(lib.dv_dpmpp_update(U.vp(out), U.vp(x), U.vp(m1c), U.vp(m0), a, b, c, 1, alp, sig, n, U.stream()))
    assert torch.equal(m1c, rt)
    assert U.rel_l2(out, a * x + b * rt + c * m0) < 1e-6
    # UniPC predictor / corrector
    xp, base = torch.empty_like(x), torch.empty_like(x)
    cp, cc, d = 0.11, -0.05, 0.4
    _lib.check(lib.dv_unipc_predict(U.vp(xp), U.vp(base), U.vp(x), U.vp(m1), U.vp(m0), a, b, cp, cc, 1, n, U.stream()))
    xbar = a * x + b * m1
    assert U.rel_l2(xp, xbar + cp * (m0 - m1)) < 1e-6 and U.rel_l2(base, xbar + cc * (m0 - m1)) < 1e-6
    _lib.check(lib.dv_unipc_correct(U.vp(out), U.vp(base), U.vp(mt), U.vp(m1), d, n, U.stream()))
    assert U.rel_l2(out, base + d * (mt - m1)) < 1e-6
    # in-place predictor without history (last UniPC step)
    xin = x.clone()
    _lib.check(lib.dv_unipc_predict(U.vp(xin), None, U.vp(xin), U.vp(m1), U.vp(m0), a, b, cp, cc, 0, n, U.stream()))
    assert U.rel_l2(xin, xbar) < 1e-6


def test_nct_to_ntc(U):
    from diffvits_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(12)
    B, Cc, T, ld = 2, 100, 203, 256
    src = torch.randn(B, Cc, T, generator=g, device="cuda")
    dst = torch.zeros(B, T, ld, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.dv_nct_to_ntc(U.vp(src), U.vp(dst), B, Cc, T, ld, 100, U.stream()))
    assert torch.equal(dst[:, :, 100:200], src.transpose(1, 2).to(torch.bfloat16))
    assert float(dst[:, :, :100].abs().max()) == 0 and float(dst[:, :, 200:].abs().max()) == 0
