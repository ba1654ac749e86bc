"""CPU-only checks (run with -m "not gpu"): the C ABI surface, the plan builder's manifest and error behaviour, the
weight-packing recipes, the host-side coefficient tables, the drop-in API shells and the multi-process plumbing."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT
from oracle import inputs, sampler_oracle


@pytest.fixture(scope="module")
def lib():
    from diffvits_b200 import _lib
    return _lib.load()


def _cfg(boc=(128, 256, 384, 512), heads=8, in_ch=228, out_ch=100):
    from diffvits_b200._lib import UnetCfg
    arr = (C.c_int * 8)(*(list(boc) + [0] * (8 - len(boc))))
    return UnetCfg(in_ch, out_ch, len(boc), arr, 2, 8, 128, heads, 1e-5, 100)


def test_abi_exports_every_declared_symbol(lib):
    from diffvits_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "diffvits_b200.h")).read()
    declared = set(re.findall(r"DV_API[^;]*?\b(dv_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.dv_version() == 103
    # measurement hooks live in their own header and only in the debug build: the product library must not export them
    dbg = open(os.path.join(ROOT, "include", "diffvits_b200_debug.h")).read()
    dbg_declared = set(re.findall(r"DV_API[^;]*?\b(dv_[a-z0-9_]+)\s*\(", dbg))
    assert dbg_declared == set(_lib.DEBUG_SIGNATURES)
    if not _lib.DEBUG_LIB:
        for name in dbg_declared:
            assert not hasattr(lib, name), name


def test_oracle_and_product_agree_on_the_workload_definition():
    from diffvits_b200 import synthetic
    from oracle import ref_import
    assert ref_import.UNET_KW == synthetic.UNET_KW


def test_compute_entry_points_fail_loudly_without_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diffvits_b200 import _lib
    rc = lib.dv_layernorm(None, None, None, None, 8, 128, 1e-5, None)
    assert rc < 0 and b"CUDA" in lib.dv_last_error() or b"device" in lib.dv_last_error()
    with pytest.raises(_lib.DVError):
        _lib.check(rc)


def test_plan_manifest_and_errors(lib, seeded_unet):
    p = C.c_void_p()
    cfg = _cfg()
    assert lib.dv_plan_create(C.byref(cfg), 2, 200, 282, C.byref(p)) == 0
    man = json.loads(lib.dv_plan_manifest_json(p).decode())
    keys = {k for s in man for k in s["keys"]}
    sd = seeded_unet.state_dict()
    missing = set(sd) - keys
    assert all(k.startswith("add_embedding.") for k in missing), missing      # hoisted prompt-only module (13 tensors)
    assert len(missing) == 13 and keys <= set(sd)
    offs = sorted((s["offset"], s["nbytes"]) for s in man)
    assert all(a + n <= b for (a, n), (b, _) in zip(offs, offs[1:]))          # slots do not overlap
    assert offs[-1][0] + offs[-1][1] <= lib.dv_plan_param_bytes(p)
    assert lib.dv_plan_op_count(p) + 3 == lib.dv_plan_launches_per_forward(p)
    assert lib.dv_plan_workspace_bytes(p) > 0 and lib.dv_plan_tensor_count(p) == 45
    lib.dv_plan_destroy(p)
    # unsupported shapes / configs report instead of crashing (reference: ValueError in the ctor)
    q = C.c_void_p()
    assert lib.dv_plan_create(C.byref(cfg), 1, 203, 50, C.byref(q)) == 0      # any length: 203 -> 102 -> 51 -> 26 frames
    assert lib.dv_plan_op_count(q) + 3 == lib.dv_plan_launches_per_forward(q)
    lib.dv_plan_destroy(q)
    assert lib.dv_plan_create(C.byref(cfg), 1, 15, 50, C.byref(q)) == -1 and b"too short" in lib.dv_last_error()
    bad = _cfg(boc=(128, 200, 384, 512))
    assert lib.dv_plan_create(C.byref(bad), 1, 200, 50, C.byref(q)) == -4
    bad = _cfg(boc=(128, 256, 384, 512))
    bad.norm_num_groups = 16                       # 128 / 16 = 8 channels per group: statistics blocks are 16 wide
    assert lib.dv_plan_create(C.byref(bad), 1, 200, 50, C.byref(q)) == -4 and b"16 * norm_num_groups" in lib.dv_last_error()
    bad = _cfg(heads=32)
    assert lib.dv_plan_create(C.byref(bad), 1, 200, 50, C.byref(q)) < 0
    assert lib.dv_plan_create(C.byref(cfg), 0, 200, 50, C.byref(q)) == -1
    assert lib.dv_plan_bind(None, None, None, None) == -1


def _tap_gemm_cpu(srcs, segs, w, l_out, views=None):
    """CPU emulation of the tap-GEMM contract (csrc/kernels.h): out[b,l,n] = sum over segments of shifted rows.
    views[si] = (rows, cols): the source's memory seen as `rows` rows of `cols` elements per batch item (batch stride
    stays L*C, as in the TMA maps); rows outside [0, rows) are zero (TMA out-of-bounds fill), anything read past the end of
    the tensor is NaN so that a view that touches memory it does not own poisons the result."""
    B = srcs[0].shape[0]
    cols = []
    for (si, ro, c0, ch) in segs:
        x = srcs[si]
        L, C = x.shape[1], x.shape[2]
        rows, vc = views[si] if views is not None and views[si] is not None else (L, C)
        flat = torch.cat([x.reshape(-1), torch.full((2 * vc,), float("nan"))])
        out = torch.zeros(B, l_out, ch)
        for b in range(B):
            for l in range(l_out):
                r = l + ro
                if 0 <= r < rows:
                    s0 = b * L * C + r * vc + c0
                    out[b, l] = flat[s0:s0 + ch]
        cols.append(out)
    a = torch.cat(cols, dim=2)
    return a @ w.t()


def test_packing_recipes_match_reference_ops():
    from diffvits_b200.engine import pack_slot
    g = torch.Generator().manual_seed(0)
    B, L, Cin, Cout = 2, 24, 64, 128
    x = torch.randn(B, L, Cin, generator=g)
    w = torch.randn(Cout, Cin, 3, generator=g)
    wsc = torch.randn(Cout, 192, 1, generator=g)
    xs = torch.randn(B, L, 192, generator=g)
    # conv3 (+ fused 1x1 shortcut over a second tensor)
    slot = dict(recipe="conv3", shape=[128, 3 * 64 + 192, 64], keys=["w", "s"], dtype="f32", name="t")
    wp = pack_slot(slot, {"w": w, "s": wsc}).reshape(128, -1)
    got = _tap_gemm_cpu([x, xs], [(0, -1, 0, 64), (0, 0, 0, 64), (0, 1, 0, 64), (1, 0, 0, 192)], wp, L)
    ref = (F.conv1d(x.transpose(1, 2), w, padding=1) + F.conv1d(xs.transpose(1, 2), wsc)).transpose(1, 2)
    assert torch.allclose(got, ref, atol=1e-4)
    # conv_in style channel padding 228 -> 256
    w_in = torch.randn(128, 228, 3, generator=g)
    x_in = torch.zeros(B, L, 256)
    x_in[:, :, :228] = torch.randn(B, L, 228, generator=g)
    wp = pack_slot(dict(recipe="conv3", shape=[128, 768, 228], keys=["w"], dtype="f32", name="t"), {"w": w_in}).reshape(128, -1)
    got = _tap_gemm_cpu([x_in], [(0, -1, 0, 256), (0, 0, 0, 256), (0, 1, 0, 256)], wp, L)
    ref = F.conv1d(x_in[:, :, :228].transpose(1, 2), w_in, padding=1).transpose(1, 2)
    assert torch.allclose(got, ref, atol=1e-4)
    # stride-2 conv through the pair view [L/2, 2C]: first-half tap over ceil(L/2) rows, second-half taps over floor(L/2)
    wd = torch.randn(64, 64, 3, generator=g)
    wp = pack_slot(dict(recipe="conv3", shape=[128, 192, 64], keys=["w"], dtype="f32", name="t"), {"w": wd}).reshape(128, -1)
    for Ld in (L, L - 1):                                      # even and odd input length
        xd = x[:, :Ld].contiguous()
        Lo = (Ld + 1) // 2
        got = _tap_gemm_cpu([xd, xd], [(1, -1, 64, 64), (0, 0, 0, 64), (1, 0, 64, 64)], wp, Lo,
                            views=[(Lo, 128), (Ld // 2, 128)])[:, :, :64]
        ref = F.conv1d(xd.transpose(1, 2), wd, stride=2, padding=1).transpose(1, 2)
        assert ref.shape[1] == Lo and torch.allclose(got, ref, atol=1e-4)
    # nearest resize to 2L (3 K segments) or 2L - 1 rows (4 segments, x' = x without its last row) folded into the weights
    wu = torch.randn(64, 64, 3, generator=g)
    wp = pack_slot(dict(recipe="conv3_up2", shape=[128, 192], keys=["w"], dtype="f32", name="t"), {"w": wu}).reshape(128, -1)
    got = _tap_gemm_cpu([x], [(0, -1, 0, 64), (0, 0, 0, 64), (0, 1, 0, 64)], wp, L).reshape(B, 2 * L, 64)
    ref = F.conv1d(F.interpolate(x.transpose(1, 2), size=2 * L, mode="nearest"), wu, padding=1).transpose(1, 2)
    assert torch.allclose(got, ref, atol=1e-4)
    wp = pack_slot(dict(recipe="conv3_up2_odd", shape=[128, 256], keys=["w"], dtype="f32", name="t"), {"w": wu}).reshape(128, -1)
    got = _tap_gemm_cpu([x, x], [(0, -1, 0, 64), (0, 0, 0, 64), (1, 0, 0, 64), (0, 1, 0, 64)], wp, L,
                        views=[None, (L - 1, 64)]).reshape(B, 2 * L, 64)[:, :2 * L - 1]
    ref = F.conv1d(F.interpolate(x.transpose(1, 2), size=2 * L - 1, mode="nearest"), wu, padding=1).transpose(1, 2)
    assert torch.allclose(got, ref, atol=1e-4)
    # GEGLU tile interleave (64 value rows, 64 gate rows, ...)
    wg = torch.randn(1024, 32, generator=g)
    wp = pack_slot(dict(recipe="geglu", shape=[1024, 32], keys=["w"], dtype="f32", name="t"), {"w": wg}).reshape(1024, 32)
    assert torch.equal(wp[0:64], wg[0:64]) and torch.equal(wp[64:128], wg[512:576]) and torch.equal(wp[128:192], wg[64:128])
    # transposed stacked time_emb_proj
    a, b = torch.randn(6, 4, generator=g), torch.randn(10, 4, generator=g)
    wt = pack_slot(dict(recipe="rows_t", shape=[4, 16], keys=["a", "b"], dtype="f32", name="t"), {"a": a, "b": b})
    assert torch.equal(wt.reshape(4, 16), torch.cat([a, b]).t())


def _fake(x, t_in, **kw):
    return torch.tanh(x * 0.7) * torch.cos(t_in * 0.01).view(-1, 1, 1) + 0.1 * torch.sin(3 * x)


def test_coefficient_tables_reproduce_the_oracle_loops():
    """The native loop's arithmetic (x = A x + B m1 + C m0 / UniPC predictor-corrector with table coefficients) against
    the independent oracle restatement, smooth model, fp32 gate 1e-4 (observed ~1e-6)."""
    from diffvits_b200.schedule import NoiseScheduleVP, dpmpp_multistep_table, get_time_steps, unipc_multistep_table
    betas = inputs.betas_linear()
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(3, 5, 16, generator=g)
    for steps in (6, 9, 10, 50):
        ns = NoiseScheduleVP("discrete", betas=betas, clip_lambda=-5.1)
        ts = get_time_steps(ns, "time_uniform", 1.0, 1e-3, steps, "cpu")
        t_in, coef = dpmpp_multistep_table(ns, ts, 2, True)
        x = x0.clone()
        m1, m0 = _fake(x, t_in[0].expand(3)), None
        for i in range(steps):
            c = coef[i]
            x = c[0] * x + c[1] * m1 + (c[2] * m0 if c[3] else 0)
            if i < steps - 1:
                m0, m1 = m1, _fake(x, t_in[i + 1].expand(3))
        want = sampler_oracle.dpmpp_2m(_fake, betas, x0.clone(), steps, roundtrip=False)
        assert float((x - want).norm() / want.norm()) < 1e-5
    for steps in (2, 3, 10, 25):
        ns = NoiseScheduleVP("discrete", betas=betas)
        ts = get_time_steps(ns, "time_uniform", 1.0, 1e-3, steps, "cpu")
        t_in, coef = unipc_multistep_table(ns, ts, 2, True, "bh2")
        x = x0.clone()
        m1, m0 = _fake(x, t_in[0].expand(3)), torch.zeros_like(x)
        for i in range(steps):
            a, b, cp, cc, d, hp, uc = coef[i][:7]
            xbar = a * x + b * m1
            dd = (m0 - m1) if hp else 0
            xp, base = xbar + cp * dd, xbar + cc * dd
            if uc:
                mt = _fake(xp, t_in[i + 1].expand(3))
                x = base + d * (mt - m1)
                m0, m1 = m1, mt
            else:
                x = xp
        want = sampler_oracle.unipc_bh2(_fake, betas, x0.clone(), steps, roundtrip=False)
        assert float((x - want).norm() / want.norm()) < 1e-5


def test_dropin_sampler_generic_loops_match_oracle(dropin):
    """Drop-in DPM_Solver / UniPC classes (reference call sequence, generic per-step path on CPU) vs the oracle,
    including batch > 1 for UniPC where the reference itself raises (uni_pc.py:189-191)."""
    dpm, uni = dropin["dpm"], dropin["uni"]
    betas = inputs.betas_linear()
    g = torch.Generator().manual_seed(2)
    x0 = torch.randn(3, 5, 16, generator=g)
    ns = dpm.NoiseScheduleVP("discrete", betas=betas)
    fn = dpm.model_wrapper(_fake, ns, model_type="x_start", model_kwargs={})
    for steps in (8, 20):
        got = dpm.DPM_Solver(fn, ns, algorithm_type="dpmsolver++").sample(x0.clone(), steps=steps, order=2,
                                                                          skip_type="time_uniform", method="multistep")
        want = sampler_oracle.dpmpp_2m(_fake, betas, x0.clone(), steps, roundtrip=True)
        assert float((got - want).norm() / want.norm()) < 1e-5
    ns2 = uni.NoiseScheduleVP("discrete", betas=betas)
    fn2 = uni.model_wrapper(_fake, ns2, model_type="x_start", model_kwargs={})
    for steps in (3, 10):
        got = uni.UniPC(fn2, ns2, variant="bh2").sample(x0.clone(), steps=steps, order=2, skip_type="time_uniform",
                                                        method="multistep")
        want = sampler_oracle.unipc_bh2(_fake, betas, x0.clone(), steps, roundtrip=True)
        assert float((got - want).norm() / want.norm()) < 1e-5
    single = dpm.DPM_Solver(fn, ns).sample(x0, steps=4, method="singlestep")      # generic loop (test_samplers_generic.py)
    assert torch.isfinite(single).all()
    with pytest.raises(ValueError):
        dpm.DPM_Solver(fn, ns).sample(x0, steps=4, method="bogus")
    with pytest.raises(ValueError):
        dpm.NoiseScheduleVP("cosine")
    assert uni.NoiseScheduleVP("cosine").T == 0.9946


def test_unet_shell_api(dropin, seeded_unet):
    U = dropin["unet"]
    with pytest.raises(ValueError):
        U.UNet1DConditionModel(block_out_channels=(128, 256), down_block_types=("DownBlock2D",) * 2,
                               up_block_types=("UpBlock2D",) * 2, resnet_time_scale_shift="scale_shift",
                               cross_attention_dim=128)
    with pytest.raises(ValueError):
        U.UNet1DConditionModel(block_out_channels=(128, 256, 384, 512))           # default 'default' scale shift
    with pytest.raises(TypeError):
        U.UNet1DConditionModel(not_an_argument=1)
    assert seeded_unet.config["block_out_channels"] == (128, 256, 384, 512)
    assert seeded_unet.config["addition_embed_type_num_heads"] == 64
    # no CPU execution path: forward refuses instead of falling back to PyTorch
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        seeded_unet(torch.zeros(1, 228, 8), 1.0, torch.zeros(1, 4, 128))
    # inference only: a training-mode call with autograd on must not silently return a detached result
    seeded_unet.train()
    try:
        with pytest.raises(RuntimeError, match="inference-only"):
            seeded_unet(torch.zeros(1, 228, 8), 1.0, torch.zeros(1, 4, 128))
    finally:
        seeded_unet.eval()
    with pytest.raises(RuntimeError, match="inference-only"):
        seeded_unet(torch.zeros(1, 228, 8, requires_grad=True), 1.0, torch.zeros(1, 4, 128))
    # state_dict round trip (what tts_infer.py:78 does)
    m2 = U.UNet1DConditionModel(**{k: seeded_unet.config[k] for k in
                                   ("in_channels", "out_channels", "block_out_channels", "norm_num_groups",
                                    "cross_attention_dim", "attention_head_dim", "addition_embed_type",
                                    "resnet_time_scale_shift")})
    assert m2.load_state_dict(seeded_unet.state_dict()).missing_keys == []
    # copy.deepcopy re-creates the module through __new__ without arguments: must not try to delegate to a reference
    # checkout (there is none on the GPU box)
    import copy
    import sys
    saved = list(sys.path)
    try:
        sys.path[:] = [e for e in sys.path if "reference" not in e]
        m3 = copy.deepcopy(seeded_unet)
    finally:
        sys.path[:] = saved
    assert type(m3) is U.UNet1DConditionModel and m3.config == seeded_unet.config


def test_find_native_unwraps_reference_closure(dropin, seeded_unet):
    """model3.py:1130-1136 wraps `self.sample_fun` in a closure; the solvers must find the Diffusion_Encoder behind it
    (here on CPU, where the native path is unavailable, detection returns None so the generic loop is used)."""
    from diffvits_b200.native import find_native

    class FakeDiff:
        unet = seeded_unet
        prompt_encoder = staticmethod(lambda p, l: p[:, :128])

    class FakeOwner:
        diff_model = FakeDiff()

        def sample_fun(self, x, t, data=None):
            return x

    owner = FakeOwner()

    def my_wrapper(fn):
        def wrapped(x, t, **kwargs):
            return fn(x, t, **kwargs)
        return wrapped

    data = (torch.zeros(1, 128, 8), torch.zeros(1, 100, 4), torch.tensor([8]), torch.tensor([4]))
    assert find_native(my_wrapper(owner.sample_fun), {"data": data}) is None       # unet lives on the CPU
    assert find_native(lambda x, t: x, {}) is None


@pytest.mark.skipif(not os.path.isfile("/root/reference/model3.py"), reason="reference checkout not present (GPU box)")
def test_reference_model3_builds_on_the_dropin():
    """`import model3; NaturalSpeech2(cfg)` with the drop-in installed: the diffusion denoiser is the native class, the
    duration predictor's small UNet (outside the accelerated configuration) is delegated to the reference class."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "helpers", "model3_dropin_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.strip().splitlines()
    assert lines[-2] == "unet1d.unet_1d_condition"
    assert "_dv_reference_unet1d.unet_1d_condition" in lines[-1] and "'unet1d.unet_1d_condition'" in lines[-1]


def test_reference_arm_of_bench_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "1", "--warmup", "0", "--cpu-target", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert len(out.stdout.strip().splitlines()) == 1          # stdout carries exactly the JSON line
    line = json.loads(out.stdout.strip())
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference"
    assert "PromptEncoder" in line["cpu_baseline"]["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def _dist_worker(rank, world, port, q):
    import torch.distributed as dist
    from diffvits_b200.dist import gather_latents, shard_inputs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    betas = inputs.betas_linear()
    x_all, cond_all, _ = inputs.make_inputs(5, 16, 4, seed=3, x_ch=6, cond_ch=6)          # ragged: 5 over 2 ranks
    x, cond = shard_inputs((x_all, cond_all), rank, world)
    lat = sampler_oracle.dpmpp_2m(lambda xx, t: _fake(xx, t) + 0.1 * cond, betas, x.clone(), 6, roundtrip=False)
    full = gather_latents(lat, 5)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sampling_equals_single_process():
    """world_size-2 gloo run of the batch-sharded loop + latent all-gather == the unsharded result, bit for bit."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    betas = inputs.betas_linear()
    x_all, cond_all, _ = inputs.make_inputs(5, 16, 4, seed=3, x_ch=6, cond_ch=6)
    want = sampler_oracle.dpmpp_2m(lambda xx, t: _fake(xx, t) + 0.1 * cond_all, betas, x_all.clone(), 6, roundtrip=False)
    assert torch.equal(full, want)
    from diffvits_b200.dist import shard_slice
    assert [shard_slice(5, r, 2) for r in range(2)] == [slice(0, 3), slice(3, 5)]


def test_serving_front_buckets_requests_faithfully():
    """Host logic of the batched front (diff-vits_b200/serving.py; reference: tts_infer.py:46-74 is batch 1): requests
    share a batch only with requests of exactly the same frame count; prompts are padded to 64-row buckets; order kept."""
    from diffvits_b200.serving import plan_batches, prompt_bucket
    assert [prompt_bucket(t) for t in (1, 64, 65, 282)] == [64, 64, 128, 320]
    shapes = [(200, 282), (136, 40), (200, 300), (136, 64), (200, 282), (136, 65), (200, 257)]
    plan = plan_batches(shapes, max_batch=2)
    assert plan == [((136, 64), [1, 3]), ((136, 128), [5]), ((200, 320), [0, 2]), ((200, 320), [4, 6])]
    seen = sorted(i for _, idx in plan for i in idx)
    assert seen == list(range(len(shapes)))
    for (T, Tp), idx in plan:
        assert all(shapes[i][0] == T and shapes[i][1] <= Tp for i in idx) and len(idx) <= 2


def test_prompt_encoder_plan_and_packing(lib):
    """Plan of the native PromptEncoder (csrc/prompt_encoder.cu): manifest keys = the reference module's state_dict keys
    (model3.py:382-433), and the 9-tap conv feed-forward packing (tap order incl. the reference's unshifted tap 0,
    operations.py:678-684, with 9^-1/2 folded in) against the reference's expression evaluated with torch ops."""
    from diffvits_b200._lib import PromptCfg
    from diffvits_b200.engine import pack_slot
    p = C.c_void_p()
    cfg = PromptCfg(100, 128, 4, 8, 9, 4)
    assert lib.dv_prompt_plan_create(C.byref(cfg), 3, 200, C.byref(p)) == 0
    man = json.loads(lib.dv_prompt_plan_manifest_json(p).decode())
    keys = {k for s in man for k in s["keys"]}
    want = {"pre.layer_norm.weight", "pre.layer_norm.bias", "pre.conv.weight", "pre.conv.bias", "layer_norm.weight",
            "layer_norm.bias", "out_proj.layer_norm.weight", "out_proj.layer_norm.bias", "out_proj.conv.weight", "out_proj.conv.bias"}
    for l in range(4):
        q = f"layers.{l}.op."
        want |= {q + "layer_norm1.weight", q + "layer_norm1.bias", q + "layer_norm2.weight", q + "layer_norm2.bias",
                 q + "self_attn.in_proj_weight", q + "self_attn.out_proj.weight", q + "ffn.ffn_1.0.bias", q + "ffn.ffn_2.weight",
                 q + "ffn.ffn_2.bias"} | {q + f"ffn.ffn_1.{i}.weight" for i in range(9)}
    assert keys == want, keys ^ want
    assert lib.dv_prompt_plan_workspace_bytes(p) > 0 and lib.dv_prompt_plan_param_bytes(p) > 0
    lib.dv_prompt_plan_destroy(p)
    bad = PromptCfg(100, 256, 4, 8, 9, 4)
    assert lib.dv_prompt_plan_create(C.byref(bad), 1, 50, C.byref(p)) == -4
    # conv feed-forward: reference expression vs tap-GEMM over the packed weights
    g = torch.Generator().manual_seed(3)
    T, B, Cc, Fh = 21, 2, 128, 512
    x = torch.randn(T, B, Cc, generator=g)
    ws = [torch.randn(Fh, Cc, generator=g) * 0.05 for _ in range(9)]
    b0 = torch.randn(Fh, generator=g)
    padded = F.pad(x, (0, 0, 0, 0, 4, 4))
    ref = sum(F.linear(padded[i:T + i] if i else x, ws[i], b0 if i == 0 else None) for i in range(9)) * 9 ** -0.5
    ref = F.relu(ref).transpose(0, 1)                                     # [B, T, F]
    slot = next(s for s in man if s["name"] == "layers.0.op.ffn.ffn_1.W")
    sd = {f"layers.0.op.ffn.ffn_1.{i}.weight": ws[i] for i in range(9)}
    sd["layers.0.op.ffn.ffn_1.0.bias"] = b0
    wp = pack_slot(slot, sd).float().reshape(slot["shape"])
    bp = pack_slot(next(s for s in man if s["name"] == "layers.0.op.ffn.ffn_1.b"), sd)
    xb = x.transpose(0, 1).contiguous()                                   # channel-last [B, T, C]
    taps = [0, -3, -2, -1, 0, 1, 2, 3, 4]
    got = F.relu(_tap_gemm_cpu([xb], [(0, ro, 0, Cc) for ro in taps], wp, T) + bp)
    assert torch.allclose(got, ref, atol=2e-2, rtol=2e-2)                 # packed weights are bf16
    # ConvTBC k = 1 with K padded from 100 to 128
    w1 = torch.randn(1, 100, 128, generator=g)
    wp1 = pack_slot(next(s for s in man if s["name"] == "pre.conv.W"), {"pre.conv.weight": w1}).float().reshape(128, 128)
    assert torch.allclose(wp1[:, :100], w1[0].t().to(torch.bfloat16).float()) and bool((wp1[:, 100:] == 0).all())
