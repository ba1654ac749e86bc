"""End-to-end parity of the native denoiser and sampling loops on the B200.

Targets
  * golden vectors produced by the REFERENCE (fp32, CPU) in tests/golden - the BF16 engine is compared with the
    stated BF16 gate of BASELINE.json where reachable and otherwise with the reference's own BF16 floor
    (reference-under-autocast-bf16 vs reference-fp32 = 1.05e-2 rel-L2 / 2e-2 max-abs per call,
    2.9e-2 / 4.8e-2 on the free-running 50-step latent): per call rel-L2 <= 1.5e-2, final latent <= 4e-2;
  * the BF16-emulating oracle (same operand rounding points, fp32 math). A 250-layer random-init network amplifies
    single rounding flips (GroupNorm couples every element of a group), so two BF16 evaluations that differ only
    in accumulation order decorrelate to the BF16 noise floor (~8e-3): gate 1.5e-2 here; the tight (<= 4e-3)
    comparisons are the per-kernel tests in test_gpu_kernels.py and the shallow-depth profile below;
  * size-independent properties at the benchmark size (batch consistency, determinism, graph == eager enqueue).
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(seeded_unet):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import gpu_util
    from oracle import inputs, sampler_oracle, unet_oracle
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    import copy
    unet = copy.deepcopy(seeded_unet).to("cuda")
    sd_gpu = {k: v.detach().clone() for k, v in unet.state_dict().items()}
    return dict(U=gpu_util, unet=unet, sd=sd_gpu, inputs=inputs, so=sampler_oracle, uo=unet_oracle)


def test_forward_c1_vs_reference_golden_and_emulated_oracle(env):
    U, unet, uo = env["U"], env["unet"], env["uo"]
    g = np.load(os.path.join(GOLD, "unet_c1.npz"))
    x_T, cond, ehs = env["inputs"].make_inputs(1, 200, 282)
    sample = torch.cat([x_T, cond], 1).cuda()
    ehs = ehs.cuda()
    for t, ref in zip(g["t"], g["out"]):
        out = unet(sample, torch.tensor([float(t)], device="cuda"), ehs).sample
        ref = torch.from_numpy(ref).cuda()
        emu = uo.unet_forward(env["sd"], sample, torch.tensor([float(t)], device="cuda"), ehs, emulate_bf16=True)
        print(f"t={t}: vs fp32 reference rel {U.rel_l2(out, ref):.3e} max {U.max_abs(out, ref):.3e} | "
              f"vs bf16-emulating oracle rel {U.rel_l2(out, emu):.3e} max {U.max_abs(out, emu):.3e} | "
              f"emulation vs fp32 rel {U.rel_l2(emu, ref):.3e}")
        assert torch.isfinite(out).all()
        assert U.rel_l2(out, emu) < 1.5e-2
        assert U.rel_l2(out, ref) < 1.2e-2 and U.max_abs(out, ref) < 2e-2


def test_forward_ragged_prompt_mask(env):
    U, unet = env["U"], env["unet"]
    g = np.load(os.path.join(GOLD, "unet_ragged.npz"))
    x, c, e = env["inputs"].make_inputs(2, 64, 282, seed=1)
    lens, mask = env["inputs"].ragged_mask(2, 282)
    e = (e * mask[:, :, None]).cuda()
    out = unet(torch.cat([x, c], 1).cuda(), torch.tensor([250.25, 250.25], device="cuda"), e,
               encoder_attention_mask=mask.cuda()).sample
    ref = torch.from_numpy(g["out"]).cuda()
    print(f"ragged: rel {U.rel_l2(out, ref):.3e} max {U.max_abs(out, ref):.3e}")
    assert U.rel_l2(out, ref) < 1.2e-2 and U.max_abs(out, ref) < 2e-2


@pytest.mark.parametrize("T,seed", [(203, 3), (198, 4)])
def test_forward_odd_lengths_vs_reference_golden(env, T, seed):
    """Any frame count, not only multiples of 8 (tts_infer.py takes T from the duration predictor): 203 -> 102 -> 51 -> 26
    and 198 -> 99 -> 50 -> 25 exercise the stride-2 conv over odd lengths (pair view with a shorter second-half map) and the
    nearest resize to 2L - 1 rows (fourth K segment over the shortened map, trimmed last half row, GroupNorm partials that
    skip it). Golden = the reference itself (oracle/make_golden.py section 3b)."""
    U, unet, uo = env["U"], env["unet"], env["uo"]
    g = np.load(os.path.join(GOLD, "unet_oddlen.npz"))
    x, c, e = env["inputs"].make_inputs(1, T, 40, seed=seed)
    sample, e = torch.cat([x, c], 1).cuda(), e.cuda()
    t = torch.tensor([613.0], device="cuda")
    out = unet(sample, t, e).sample
    ref = torch.from_numpy(g[f"out{T}"]).cuda()
    emu = uo.unet_forward(env["sd"], sample, t, e, emulate_bf16=True)
    print(f"T={T}: vs fp32 reference rel {U.rel_l2(out, ref):.3e} max {U.max_abs(out, ref):.3e} | "
          f"vs bf16-emulating oracle rel {U.rel_l2(out, emu):.3e}")
    assert out.shape == ref.shape and torch.isfinite(out).all()
    assert U.rel_l2(out, emu) < 1.5e-2
    assert U.rel_l2(out, ref) < 1.2e-2 and U.max_abs(out, ref) < 2e-2
    # batch of 3 at the same odd length: every sample equals its single-sample evaluation (batch strides, trimmed rows).
    # The hoisted torch op (add_embedding, cuBLAS) is shared: it picks batch-size dependent kernels.
    import diffvits_b200
    xb, cb, eb = (v.cuda() for v in env["inputs"].make_inputs(3, T, 40, seed=seed + 10))
    with torch.no_grad():
        aug = unet.add_embedding(eb)
    full = diffvits_b200.NativeDenoiser(unet, cb, eb, aug=aug)(xb, t).clone()
    assert torch.isfinite(full).all()
    for i in range(3):
        one = diffvits_b200.NativeDenoiser(unet, cb[i:i + 1].contiguous(), eb[i:i + 1].contiguous(), aug=aug[i:i + 1])
        assert U.rel_l2(full[i:i + 1], one(xb[i:i + 1].contiguous(), t)) < 1e-6, i


def test_sampling_loop_odd_length_vs_oracle_loop(env):
    """Whole native DPM-Solver++ loop (graph) at T = 203 against the oracle loop driven by the same native denoiser.
    Gate = the BF16 floor (see test_native_loop_vs_oracle_loop_with_shared_denoiser): a 1-ulp difference in x flips operand
    roundings of the network; observed either ~1e-8 or ~6e-3 depending on T and the step count."""
    import diffvits_b200
    U, unet, so = env["U"], env["unet"], env["so"]
    x, c, e = (v.cuda() for v in env["inputs"].make_inputs(2, 203, 40, seed=5))
    betas = env["inputs"].betas_linear()
    native = diffvits_b200.NativeDenoiser(unet, c, e)
    lat = diffvits_b200.sample_native(unet, x, c, e, sampler="dpmsolver++", steps=8, betas=betas, native=native)
    want = so.dpmpp_2m(lambda xx, ti: native(xx, ti.to("cuda")), betas, x.clone(), 8, roundtrip=False)
    print(f"T=203 8-step loop vs oracle loop rel {U.rel_l2(lat, want):.3e}")
    assert torch.isfinite(lat).all() and U.rel_l2(lat, want) < 4e-2


def test_forward_api_variants(env):
    """timestep as python float / 0-d tensor / [B] tensor give the same result; return_dict=False returns a tuple."""
    unet = env["unet"]
    x, c, e = env["inputs"].make_inputs(2, 64, 40, seed=2)
    s, e = torch.cat([x, c], 1).cuda(), e.cuda()
    a = unet(s, 123.5, e).sample
    b = unet(s, torch.tensor(123.5), e).sample
    d = unet(s, torch.tensor([123.5, 123.5], device="cuda"), e, return_dict=False)[0]
    assert torch.equal(a, b) and torch.equal(a, d)
    with pytest.raises(NotImplementedError):
        unet(s, 1.0, e, class_labels=torch.zeros(2))


def test_sampling_loops_vs_reference_golden(env):
    """Free-running native loops (CUDA graph) from the same noise vs the reference's final latents, plus
    teacher-forced per-step denoiser outputs at the reference's trajectory points."""
    import diffvits_b200
    U, unet = env["U"], env["unet"]
    g = np.load(os.path.join(GOLD, "sampler_c1.npz"))
    x_T, cond, ehs = env["inputs"].make_inputs(1, 200, 282)
    betas = env["inputs"].betas_linear()
    native = diffvits_b200.NativeDenoiser(unet, cond.cuda(), ehs.cuda())
    for name, sampler, steps in (("dpm50", "dpmsolver++", 50), ("unipc10", "unipc", 10)):
        lat = diffvits_b200.sample_native(unet, x_T.cuda(), cond.cuda(), ehs.cuda(), sampler=sampler, steps=steps,
                                          betas=betas, native=native)
        ref = torch.from_numpy(g[name]).cuda()
        print(f"{name}: final latent rel {U.rel_l2(lat, ref):.3e} max {U.max_abs(lat, ref):.3e}")
        assert torch.isfinite(lat).all()
        assert U.rel_l2(lat, ref) < 4e-2
        eager = diffvits_b200.sample_native(unet, x_T.cuda(), cond.cuda(), ehs.cuda(), sampler=sampler, steps=steps,
                                            betas=betas, native=native, graph=False)
        assert torch.equal(lat, eager)          # graph replay == direct enqueue, bit for bit
        again = diffvits_b200.sample_native(unet, x_T.cuda(), cond.cuda(), ehs.cuda(), sampler=sampler, steps=steps,
                                            betas=betas, native=native)
        assert torch.equal(lat, again)          # second sight of the loop = captured graph replay: deterministic, same bits
    # first-step denoiser outputs are teacher forced by construction (same x_T)
    first = native(x_T.cuda(), torch.tensor([999.0], device="cuda"))
    ref0 = torch.from_numpy(g["dpm_steps"][0]).cuda()
    assert U.rel_l2(first, ref0) < 1.5e-2


def test_depth_profile_vs_emulating_oracle(env):
    """Block-by-block comparison with the BF16-emulating oracle (reference module order): the first blocks agree to a
    rounding or two, the distance then grows towards the BF16 noise floor as rounding flips are amplified by depth."""
    U, unet, uo = env["U"], env["unet"], env["uo"]
    import diffvits_b200
    x_T, cond, ehs = env["inputs"].make_inputs(2, 200, 282, seed=9)
    native = diffvits_b200.NativeDenoiser(unet, cond.cuda(), ehs.cuda())
    t = torch.tensor([400.0], device="cuda")
    native(x_T.cuda(), t)
    torch.cuda.synchronize()
    got = {k: v.float().transpose(1, 2) for k, v in native.engine.named_tensors().items()}
    trace = {}
    uo.unet_forward(env["sd"], torch.cat([x_T, cond], 1).cuda(), t, ehs.cuda(), emulate_bf16=True, trace=trace)
    assert set(got) == set(trace), set(got) ^ set(trace)
    rels = {k: U.rel_l2(got[k], trace[k]) for k in trace}
    for k in trace:
        print(f"  {k:40s} rel {rels[k]:.3e}")
    assert rels["conv_in"] < 3e-3                        # one conv: ulp-level flips only
    assert rels["down_blocks.0.resnets.0"] < 6e-3        # GN + 2 convs
    assert rels["down_blocks.0.attentions.0"] < 8e-3     # + GN, 9 GEMMs, 2 attentions, 3 LNs
    assert rels["down_blocks.0.downsamplers.0.conv"] < 1e-2
    assert max(rels.values()) < 2e-2                     # never worse than the BF16 floor anywhere in the network


def test_teacher_forced_per_block_parity(env):
    """Every block of the network on its own: the emulating oracle evaluates block k on the ENGINE's input of block k (the
    named tensors the engine leaves in its workspace), so a block's error does not compound with depth and the gate can be
    two orders of magnitude below the end-to-end BF16 floor. This is what proves the kernel logic at depth: conv taps and
    padding at every level, stride-2 / folded-upsample packing, two-source GroupNorm + shortcut over the never-materialised
    concat, time scale/shift rows, attention + key padding, GEGLU, residual epilogues - for all 45 block outputs and
    the final conv_out. What remains is accumulation order and 1-ulp flips of the BF16 intermediates INSIDE one block."""
    U, unet, uo = env["U"], env["unet"], env["uo"]
    import diffvits_b200
    x_T, cond, ehs = env["inputs"].make_inputs(2, 200, 282, seed=21)
    lens, mask = env["inputs"].ragged_mask(2, 282)
    ehs = ehs * mask[:, :, None]
    native = diffvits_b200.NativeDenoiser(unet, cond.cuda(), ehs.cuda(), kv_len=lens.cuda())
    t = torch.tensor([731.5], device="cuda")
    out = native(x_T.cuda(), t).clone()
    torch.cuda.synchronize()
    got = {k: v.float().transpose(1, 2).contiguous() for k, v in native.engine.named_tensors().items()}
    assert len(got) == 45
    trace = {}
    final = uo.unet_forward(env["sd"], torch.cat([x_T, cond], 1).cuda(), t, ehs.cuda(), mask=mask.cuda(), emulate_bf16=True,
                            trace=trace, teacher=got)
    rels = {k: U.rel_l2(got[k], trace[k]) for k in trace}
    rels["conv_out"] = U.rel_l2(out, final)
    worst = sorted(rels.items(), key=lambda kv: -kv[1])
    for k, v in worst[:8]:
        print(f"  {k:40s} rel {v:.3e}")
    res = [v for k, v in rels.items() if ".resnets." in k]
    att = [v for k, v in rels.items() if ".attentions." in k]
    print(f"teacher-forced per-block: {len(rels)} tensors, resnets max {max(res):.2e}, transformers max {max(att):.2e}, "
          f"conv_in {rels['conv_in']:.2e}, conv_out {rels['conv_out']:.2e}")
    # measured (B200, r2a): conv_in 5.6e-5, conv_out 5.1e-4, resnets <= 2.5e-3, transformer blocks <= 2.6e-3 - the BF16
    # rounding flips of the 4 (resnet) / ~20 (transformer) intermediate tensors inside ONE block, 3-4x below the end-to-end
    # floor. The FP32-class mode runs the same test at 1e-4 (tests/test_gpu_precision.py).
    assert rels["conv_in"] < 1e-3 and rels["conv_out"] < 2e-3
    for k, v in rels.items():
        assert v < 3.5e-3, (k, v)


def test_teacher_forced_per_step_parity(env):
    """Denoiser outputs at the REFERENCE's own trajectory points (steps 0 / 1 / 25 / 49 of the 50-step DPM-Solver++ loop,
    0 / 1 / 9 of the 10-step UniPC loop): the engine is fed the reference's x_k (tests/golden/sampler_c1_inputs.npz) and
    compared with the reference's output of that call, so late steps (small t, structured x) are covered without the
    trajectory drift of a free-running loop."""
    U, unet = env["U"], env["unet"]
    import diffvits_b200
    g = np.load(os.path.join(GOLD, "sampler_c1.npz"))
    gi = np.load(os.path.join(GOLD, "sampler_c1_inputs.npz"))
    x_T, cond, ehs = env["inputs"].make_inputs(1, 200, 282)
    native = diffvits_b200.NativeDenoiser(unet, cond.cuda(), ehs.cuda())
    for name, xs, ts, refs in (("dpm", gi["dpm_x"], gi["dpm_t"], g["dpm_steps"]), ("unipc", gi["uni_x"], gi["uni_t"], g["uni_steps"])):
        for x, t, ref in zip(xs, ts, refs):
            out = native(torch.from_numpy(x).cuda(), torch.tensor([float(t)], device="cuda"))
            ref = torch.from_numpy(ref).cuda()
            rel, mx = U.rel_l2(out, ref), U.max_abs(out, ref)
            print(f"{name} t_in={float(t):8.3f}: teacher-forced denoiser output vs reference rel {rel:.3e} max {mx:.3e}")
            assert torch.isfinite(out).all()
            assert mx < 2e-2 and rel < 1.2e-2      # north_star max-abs gate; rel-L2 = measured BF16 floor (8-9e-3) + margin


def test_loop_control_bit_exact(env):
    """The captured whole-loop graph equals a host-driven loop issuing the same kernels step by step through the C ABI
    (dv_unet_forward + dv_dpmpp_update / dv_unipc_*), bit for bit: checks history rotation, orders, last-step rules."""
    import diffvits_b200
    from diffvits_b200 import _lib
    from diffvits_b200.schedule import (NoiseScheduleVP, dpmpp_multistep_table, get_time_steps,
                                        unipc_multistep_table)
    U, unet = env["U"], env["unet"]
    lib = _lib.load()
    x_T, cond, ehs = env["inputs"].make_inputs(2, 64, 40, seed=4)
    betas = env["inputs"].betas_linear()
    native = diffvits_b200.NativeDenoiser(unet, cond.cuda(), ehs.cuda())
    n = x_T.numel()
    for sampler, steps in (("dpmsolver++", 12), ("dpmsolver++", 6), ("unipc", 10), ("unipc", 3)):
        ns = NoiseScheduleVP("discrete", betas=betas, clip_lambda=-5.1 if sampler == "dpmsolver++" else None)
        ts = get_time_steps(ns, "time_uniform", 1.0, 1e-3, steps, "cpu")
        x = x_T.cuda().clone()
        if sampler == "dpmsolver++":
            t_in, coef = dpmpp_multistep_table(ns, ts, 2, True)
            m1, m0 = native(x, t_in[0:1]), None
            for i in range(steps):
                a, b, c, hp = (float(v) for v in coef[i][:4])
                _lib.check(lib.dv_dpmpp_update(U.vp(x), U.vp(x), U.vp(m1), U.vp(m0) if hp else None, a, b, c, 0, 1.0,
                                               1.0, n, U.stream()))
                if i < steps - 1:
                    m0, m1 = m1, native(x, t_in[i + 1:i + 2])
        else:
            t_in, coef = unipc_multistep_table(ns, ts, 2, True)
            m1, m0 = native(x, t_in[0:1]), torch.zeros_like(x)
            xp, base = torch.empty_like(x), torch.empty_like(x)
            for i in range(steps):
                a, b, cp, cc, d, hp, uc = (float(v) for v in coef[i][:7])
                if uc:
                    _lib.check(lib.dv_unipc_predict(U.vp(xp), U.vp(base), U.vp(x), U.vp(m1), U.vp(m0), a, b, cp, cc,
                                                    int(hp), n, U.stream()))
                    mt = native(xp, t_in[i + 1:i + 2])
                    _lib.check(lib.dv_unipc_correct(U.vp(x), U.vp(base), U.vp(mt), U.vp(m1), d, n, U.stream()))
                    m0, m1 = m1, mt
                else:
                    _lib.check(lib.dv_unipc_predict(U.vp(x), None, U.vp(x), U.vp(m1), U.vp(m0), a, b, cp, cc, int(hp),
                                                    n, U.stream()))
        got = diffvits_b200.sample_native(unet, x_T.cuda(), cond.cuda(), ehs.cuda(), sampler=sampler, steps=steps,
                                          betas=betas, native=native)
        assert torch.equal(got, x), (sampler, steps)


def test_native_loop_vs_oracle_loop_with_shared_denoiser(env):
    """ORACLE sampler loops driven by the native denoiser vs the native loop. The coefficient arithmetic agrees to
    ~1e-7 (checked with smooth models in test_host.py); with the BF16 network in the loop a 1-ulp difference in x can
    flip operand roundings, so the gate here is the BF16 floor, not 1e-4."""
    import diffvits_b200
    U, unet, so = env["U"], env["unet"], env["so"]
    x_T, cond, ehs = env["inputs"].make_inputs(2, 64, 40, seed=4)
    betas = env["inputs"].betas_linear()
    native = diffvits_b200.NativeDenoiser(unet, cond.cuda(), ehs.cuda())

    def model(x, t_in):
        return native(x, t_in.cuda()).clone()

    for sampler, steps, fn in (("dpmsolver++", 12, so.dpmpp_2m), ("dpmsolver++", 6, so.dpmpp_2m),
                               ("unipc", 10, so.unipc_bh2), ("unipc", 3, so.unipc_bh2)):
        want = fn(model, betas, x_T.cuda().clone(), steps, roundtrip=False)
        got = diffvits_b200.sample_native(unet, x_T.cuda(), cond.cuda(), ehs.cuda(), sampler=sampler, steps=steps,
                                          betas=betas, native=native)
        print(f"{sampler}/{steps}: native loop vs oracle loop rel {U.rel_l2(got, want):.3e}")
        assert U.rel_l2(got, want) < 4e-2
    want = so.dpmpp_2m(model, betas, x_T.cuda().clone(), 12, roundtrip=True)
    got = diffvits_b200.sample_native(unet, x_T.cuda(), cond.cuda(), ehs.cuda(), sampler="dpmsolver++", steps=12,
                                      betas=betas, native=native, roundtrip=True)
    assert U.rel_l2(got, want) < 4e-2


def test_dropin_sampler_api_runs_native(env, dropin):
    """The reference call sequence of model3.py:1143-1158 / 1177-1192 on the drop-in classes takes the native path and
    equals sample_native; an unsupported setting falls back to the generic per-step loop over the same denoiser."""
    import diffvits_b200
    unet = env["unet"]
    x_T, cond, ehs = env["inputs"].make_inputs(2, 64, 40, seed=6)
    betas = env["inputs"].betas_linear()
    native = diffvits_b200.NativeDenoiser(unet, cond.cuda(), ehs.cuda())
    dpm, uni = dropin["dpm"], dropin["uni"]
    ns = dpm.NoiseScheduleVP("discrete", betas=betas)
    fn = dpm.model_wrapper(native, ns, model_type="x_start", model_kwargs={})
    solver = dpm.DPM_Solver(fn, ns, algorithm_type="dpmsolver++")
    a = solver.sample(x_T.cuda(), steps=20, order=2, skip_type="time_uniform", method="multistep")
    assert solver.last_run_native
    b = diffvits_b200.sample_native(unet, x_T.cuda(), cond.cuda(), ehs.cuda(), sampler="dpmsolver++", steps=20,
                                    betas=betas, native=native)
    assert torch.equal(a, b)
    c = solver.sample(x_T.cuda(), steps=20, order=3, skip_type="time_uniform", method="multistep")   # generic loop
    assert not solver.last_run_native and torch.isfinite(c).all()
    generic2 = dpm.DPM_Solver(dpm.model_wrapper(lambda x, t, **k: native(x, t), ns, model_type="x_start"), ns,
                              "dpmsolver++").sample(x_T.cuda(), steps=20, order=2, method="multistep")
    assert env["U"].rel_l2(generic2, a) < 4e-2       # per-step Python loop over the native forward ~ graph loop (BF16 floor)
    ns2 = uni.NoiseScheduleVP("discrete", betas=betas)
    fn2 = uni.model_wrapper(native, ns2, model_type="x_start", model_kwargs={})
    up = uni.UniPC(fn2, ns2, variant="bh2")
    d = up.sample(x_T.cuda(), steps=10, order=2, skip_type="time_uniform", method="multistep")
    assert up.last_run_native
    e = diffvits_b200.sample_native(unet, x_T.cuda(), cond.cuda(), ehs.cuda(), sampler="unipc", steps=10, betas=betas,
                                    native=native)
    assert torch.equal(d, e)


def test_sampler_diagnostics_and_nan_assertion(env, dropin):
    """Device-side sampler diagnostics (north_star item 3: reductions inside the update kernels) that stand in for the
    reference's per-call `assert torch.isnan(x).any() == False` (model3.py:903): one row per sampler-update launch with
    the non-finite count and the sums of squares of the denoiser output consumed / the state written. Checked against
    torch on a host-driven replay of the same loop, identical between graph replay and direct enqueue, and a NaN put
    into the conditioning raises the reference's AssertionError from the drop-in solver."""
    import diffvits_b200
    from diffvits_b200.engine import SAMPLER_DPMPP_2M, SAMPLER_UNIPC_BH2
    from diffvits_b200.schedule import (NoiseScheduleVP, dpmpp_multistep_table, get_time_steps, unipc_multistep_table)
    unet = env["unet"]
    x_T, cond, ehs = env["inputs"].make_inputs(2, 72, 40, seed=8)
    betas = env["inputs"].betas_linear()
    native = diffvits_b200.NativeDenoiser(unet, cond.cuda(), ehs.cuda())
    eng = native.engine
    ns = NoiseScheduleVP("discrete", betas=betas)
    for sampler, steps, want_rows in (("dpmsolver++", 12, 12), ("unipc", 6, 11)):
        ts = get_time_steps(ns, "time_uniform", ns.T, 1.0 / ns.total_N, steps, "cpu")
        if sampler == "dpmsolver++":
            t_in, coef = dpmpp_multistep_table(ns, ts, 2, True, "dpmsolver")
            kind = SAMPLER_DPMPP_2M
        else:
            t_in, coef = unipc_multistep_table(ns, ts, 2, True, "bh2")
            kind = SAMPLER_UNIPC_BH2
        diags = []
        for graph in (False, True, True):          # first sight: direct enqueue; third call: graph replay
            out = eng.sample(kind, t_in, coef, x_T.cuda().clone(), graph=graph)
            diags.append(eng.sample_diag().cpu().clone())
        d = diags[0]
        assert d.shape == (want_rows, 4), d.shape
        assert torch.equal(diags[0], diags[1]) and torch.equal(diags[0], diags[2])     # fixed-order reductions
        assert float(d[:, 0].abs().sum()) == 0.0
        # last row: the state it wrote is the returned latent
        assert abs(float(d[-1, 2]) / float((out.double() ** 2).sum()) - 1.0) < 1e-5
        if sampler == "dpmsolver++":
            assert set(d[:, 3].tolist()) == {0.0}
            # row 0 consumed the first denoiser output m0 = model(x_T, t_0)
            m0 = native(x_T.cuda(), t_in[:1].cuda())
            assert abs(float(d[0, 1]) / float((m0.double() ** 2).sum()) - 1.0) < 1e-5
        else:
            assert d[:, 3].tolist() == [1.0, 2.0] * 5 + [1.0]      # predictor + corrector per step, last step predictor only
        eng.check_finite()
    # a NaN in the conditioning: the loop runs to the end (no per-step sync), the solver raises like the reference would
    bad = cond.clone()
    bad[1, 3, 5] = float("nan")
    native.rebind(bad.cuda(), ehs.cuda())
    dpm = dropin["dpm"]
    ns2 = dpm.NoiseScheduleVP("discrete", betas=betas)
    solver = dpm.DPM_Solver(dpm.model_wrapper(native, ns2, model_type="x_start", model_kwargs={}), ns2,
                            algorithm_type="dpmsolver++")
    with pytest.raises(AssertionError, match="non-finite"):
        solver.sample(x_T.cuda(), steps=6, order=2, skip_type="time_uniform", method="multistep")
    assert float(eng.sample_diag()[0, 0]) > 0
    native.check_finite = False
    solver.sample(x_T.cuda(), steps=6, order=2, skip_type="time_uniform", method="multistep")     # opt-out: no read


def test_batch_consistency_at_bench_size(env):
    """Size-independent property at the benchmark shape (B=32, T=1000): every utterance of the batch equals the same
    utterance run alone (no cross-sample reduction exists on the path), which also covers multi-GPU sharding."""
    import diffvits_b200
    U, unet = env["U"], env["unet"]
    B, T, Tp = 32, 1000, 282
    x_T, cond, ehs = env["inputs"].make_inputs(B, T, Tp, seed=7)
    t = torch.tensor([321.0], device="cuda")
    with torch.no_grad():
        aug = unet.add_embedding(ehs.cuda())          # hoisted torch op; cuBLAS picks batch-size dependent kernels, so
    big = diffvits_b200.NativeDenoiser(unet, cond.cuda(), ehs.cuda(), aug=aug)     # share it to isolate the engine
    out = big(x_T.cuda(), t).clone()
    assert torch.isfinite(out).all()
    for b in (0, 17, 31):
        one = diffvits_b200.NativeDenoiser(unet, cond[b:b + 1].cuda(), ehs[b:b + 1].cuda(), aug=aug[b:b + 1])
        o1 = one(x_T[b:b + 1].cuda(), t)
        assert U.rel_l2(out[b:b + 1], o1) < 1e-6, b    # the hoisted cuBLAS op aside, reductions are fixed-order (no atomics)
    emu = env["uo"].unet_forward(env["sd"], torch.cat([x_T[:2], cond[:2]], 1).cuda(), t, ehs[:2].cuda(),
                                 emulate_bf16=True)
    print(f"B=32,T=1000 rows 0-1 vs emulating oracle: rel {U.rel_l2(out[:2], emu):.3e}")
    assert U.rel_l2(out[:2], emu) < 1.5e-2
