"""Parity at the BASELINE.json workload shapes that the small-size tests do not reach.

  C3: batch 32 x 1000 frames, 10-step UniPC with ragged prompt cross-attention (key-padding masks)
  C5: long-form 4000 frames (self-attention over 4000 keys at level 0), 25-step UniPC
plus the "kernel bar" of SURVEY.md section 8d: the same forward evaluated by eager PyTorch (cuDNN / cuBLAS / SDPA
through the oracle restatement, which uses the reference's own torch ops) on the same GPU, timed next to the
native call. /root/reference does not exist on the GPU box, so the oracle port stands in for the reference here.

Gates follow test_gpu_unet.py: a BF16 network evaluated with a different accumulation order decorrelates to the BF16
floor (~1e-2 rel-L2 per call, SURVEY section 0: the reference's own autocast-bf16 run is at 1.05e-2), so per-call
comparisons against the BF16-emulating oracle are gated at 1.5e-2; loop-control equivalence is checked bit for bit.
"""
import copy
import time

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(seeded_unet):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gpu_util
    from oracle import inputs, sampler_oracle, unet_oracle
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    unet = copy.deepcopy(seeded_unet).to("cuda")
    sd = {k: v.detach().clone() for k, v in unet.state_dict().items()}
    return dict(U=gpu_util, unet=unet, sd=sd, inputs=inputs, so=sampler_oracle, uo=unet_oracle)


def test_c3_ragged_prompt_call_and_unipc_loop(env):
    """C3 shape with ragged prompts: per-call parity for masked cross-attention rows, the captured 10-step UniPC loop
    equals the host-driven sequence of the same kernels bit for bit, and every utterance equals itself run alone."""
    import diffvits_b200
    from diffvits_b200 import _lib
    from diffvits_b200.schedule import NoiseScheduleVP, get_time_steps, unipc_multistep_table
    U, unet, uo = env["U"], env["unet"], env["uo"]
    lib = _lib.load()
    B, T, Tp = 32, 1000, 282
    x_T, cond, ehs = env["inputs"].make_inputs(B, T, Tp, seed=21)
    lens, mask = env["inputs"].ragged_mask(B, Tp, step=9)               # 282, 273, ... 3 valid prompt rows
    ehs = (ehs * mask[:, :, None]).cuda()
    x_T, cond = x_T.cuda(), cond.cuda()
    betas = env["inputs"].betas_linear()
    native = diffvits_b200.NativeDenoiser(unet, cond, ehs, kv_len=lens.cuda())
    t = torch.tensor([640.0], device="cuda")
    out = native(x_T, t).clone()
    assert torch.isfinite(out).all()
    rows = [0, 13, 31]                                                   # full prompt, partial, almost fully masked
    emu = uo.unet_forward(env["sd"], torch.cat([x_T[rows], cond[rows]], 1), t, ehs[rows], mask=mask[rows].cuda(),
                          emulate_bf16=True)
    for i, r in enumerate(rows):
        rel = U.rel_l2(out[r:r + 1], emu[i:i + 1])
        print(f"C3 call row {r} (prompt length {int(lens[r])}): vs emulating oracle rel {rel:.3e}")
        assert rel < 1.5e-2
    # loop control at the full size
    steps = 10
    ns = NoiseScheduleVP("discrete", betas=betas)
    ts = get_time_steps(ns, "time_uniform", 1.0, 1e-3, steps, "cpu")
    t_in, coef = unipc_multistep_table(ns, ts, 2, True)
    n = x_T.numel()
    x = x_T.clone()
    m1, m0 = native(x, t_in[0:1]).clone(), torch.zeros_like(x)
    xp, base = torch.empty_like(x), torch.empty_like(x)
    for i in range(steps):
        a, b, cp, cc, d, hp, uc = (float(v) for v in coef[i][:7])
        if uc:
            _lib.check(lib.dv_unipc_predict(U.vp(xp), U.vp(base), U.vp(x), U.vp(m1), U.vp(m0), a, b, cp, cc, int(hp), n,
                                            U.stream()))
            mt = native(xp, t_in[i + 1:i + 2]).clone()
            _lib.check(lib.dv_unipc_correct(U.vp(x), U.vp(base), U.vp(mt), U.vp(m1), d, n, U.stream()))
            m0, m1 = m1, mt
        else:
            _lib.check(lib.dv_unipc_predict(U.vp(x), None, U.vp(x), U.vp(m1), U.vp(m0), a, b, cp, cc, int(hp), n,
                                            U.stream()))
    got = diffvits_b200.sample_native(unet, x_T, cond, ehs, sampler="unipc", steps=steps, betas=betas, native=native)
    assert torch.isfinite(got).all()
    assert torch.equal(got, x)
    # an utterance of the batch == the same utterance alone (ragged masks are per utterance)
    with torch.no_grad():
        aug = unet.add_embedding(ehs)
    r = 13
    one = diffvits_b200.NativeDenoiser(unet, cond[r:r + 1], ehs[r:r + 1], kv_len=lens[r:r + 1].cuda(), aug=aug[r:r + 1])
    big = diffvits_b200.NativeDenoiser(unet, cond, ehs, kv_len=lens.cuda(), aug=aug)
    assert U.rel_l2(big(x_T, t)[r:r + 1], one(x_T[r:r + 1], t)) < 1e-6


@pytest.mark.parametrize("B,T,Tp,lens", [(1, 16, 1, None), (2, 17, 5, [5, 0]), (3, 33, 282, [282, 1, 140]),
                                         (1, 1023, 64, None), (5, 136, 40, None)])
def test_edge_shapes_vs_emulating_oracle(env, B, T, Tp, lens):
    """Smallest frame counts the four levels allow (16 -> 8 -> 4 -> 2, 17 -> 9 -> 5 -> 3), a one-frame prompt, a fully
    masked prompt (length 0: the reference's -10000 bias then cancels in the softmax and every key is attended), odd
    lengths just below a tile boundary, a batch that is not a multiple of anything. Checked against the BF16-emulating
    oracle (itself pinned to the reference at T = 200 / 203 / 198 / 64 by goldens)."""
    U, unet, uo = env["U"], env["unet"], env["uo"]
    x, c, e = (v.cuda() for v in env["inputs"].make_inputs(B, T, Tp, seed=100 + T))
    mask = None
    if lens is not None:
        mask = (torch.arange(Tp, device="cuda")[None, :] < torch.tensor(lens, device="cuda")[:, None])
        e = e * mask[:, :, None]
    t = torch.full((B,), 431.5, device="cuda")
    sample = torch.cat([x, c], 1)
    out = unet(sample, t, e, encoder_attention_mask=mask).sample
    emu = uo.unet_forward(env["sd"], sample, t, e, mask=mask, emulate_bf16=True)
    ref = uo.unet_forward(env["sd"], sample, t, e, mask=mask)
    print(f"B={B} T={T} Tp={Tp} lens={lens}: vs emulating oracle rel {U.rel_l2(out, emu):.3e}, vs fp32 oracle "
          f"{U.rel_l2(out, ref):.3e} max {U.max_abs(out, ref):.3e}")
    assert out.shape == (B, 100, T) and torch.isfinite(out).all()
    assert U.rel_l2(out, emu) < 1.5e-2 and U.rel_l2(out, ref) < 1.5e-2 and U.max_abs(out, ref) < 4e-2


def test_c5_long_form_call_and_loop(env):
    """C5 shape (T = 4000): one call against the emulating oracle, then the 25-step UniPC loop: finite, deterministic,
    graph replay == direct enqueue."""
    import diffvits_b200
    U, unet, uo = env["U"], env["unet"], env["uo"]
    B, T, Tp = 2, 4000, 282
    x_T, cond, ehs = (v.cuda() for v in env["inputs"].make_inputs(B, T, Tp, seed=22))
    betas = env["inputs"].betas_linear()
    native = diffvits_b200.NativeDenoiser(unet, cond, ehs)
    t = torch.tensor([77.0], device="cuda")
    out = native(x_T, t).clone()
    emu = uo.unet_forward(env["sd"], torch.cat([x_T[:1], cond[:1]], 1), t, ehs[:1], emulate_bf16=True)
    rel = U.rel_l2(out[:1], emu)
    print(f"C5 call (T=4000): vs emulating oracle rel {rel:.3e} max {U.max_abs(out[:1], emu):.3e}")
    assert torch.isfinite(out).all() and rel < 1.5e-2
    a = diffvits_b200.sample_native(unet, x_T, cond, ehs, sampler="unipc", steps=25, betas=betas, native=native)
    b = diffvits_b200.sample_native(unet, x_T, cond, ehs, sampler="unipc", steps=25, betas=betas, native=native,
                                    graph=False)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    # the first sight of a loop runs eagerly, the second one is captured: this one is the graph replay
    c = diffvits_b200.sample_native(unet, x_T, cond, ehs, sampler="unipc", steps=25, betas=betas, native=native)
    assert torch.equal(a, c)
    # the same 25-step loop by the ORACLE's UniPC restatement (pinned to the reference's sampler in tests/test_oracle.py)
    # driving the native denoiser: checks the loop control and coefficient tables at the C5 step count
    want = env["so"].unipc_bh2(lambda xx, ti: native(xx, ti.to("cuda")).clone(), betas, x_T.clone(), 25)
    rel = U.rel_l2(a, want)
    print(f"C5 25-step UniPC loop (T=4000): native loop vs oracle loop over the native denoiser rel {rel:.3e}")
    assert rel < 4e-2


def test_eager_pytorch_bar_on_same_gpu(env):
    """The kernel "bar": one C2 denoiser call (32 x 1000) by eager PyTorch on this GPU (fp32 with TF32 off, and
    bf16 autocast) next to the native call. The native engine must beat both; the numbers are printed for DESIGN.md."""
    import diffvits_b200
    unet, uo = env["unet"], env["uo"]
    B, T, Tp = 32, 1000, 282
    x_T, cond, ehs = (v.cuda() for v in env["inputs"].make_inputs(B, T, Tp, seed=23))
    sample = torch.cat([x_T, cond], 1)
    t = torch.tensor([500.0], device="cuda")
    native = diffvits_b200.NativeDenoiser(unet, cond, ehs)

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e3

    with torch.no_grad():
        ms_native = timed(lambda: native(x_T, t), 10)
        ms_fp32 = timed(lambda: uo.unet_forward(env["sd"], sample, t, ehs), 3)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms_bf16 = timed(lambda: uo.unet_forward(env["sd"], sample, t, ehs), 3)
    print(f"C2 denoiser call: native {ms_native:.2f} ms (eager enqueue, no graph) | eager PyTorch fp32 {ms_fp32:.2f} ms | "
          f"eager PyTorch autocast-bf16 {ms_bf16:.2f} ms")
    assert ms_native < ms_bf16 and ms_native < ms_fp32
