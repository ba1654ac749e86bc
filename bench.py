#!/usr/bin/env python
"""Benchmark of the Diff-VITS denoising loop (BASELINE.json metric: denoised latent frames/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5|c1]
                    [--impl native|reference|gpu-reference] [--scaling weak|strong] [--global-batch B]
                    [--precision bf16|fp16|fp16s|fp32x3]

One "step" = one complete sampling loop (all denoiser calls + sampler updates) over one batch of synthetic
utterances. N > 1: launched under torchrun, one rank per GPU, independent utterances batch-sharded, one all-gather of the
output latents per step, max-over-ranks device timing. `--scaling weak` (default) keeps the per-GPU batch of the workload;
`--scaling strong` (or `--workload c4`) splits a fixed global batch (C4: 256) over the ranks.
Rank 0 prints ONE JSON line. See DESIGN.md section "Measurement" for every field.

Arms:  native          the product path (libdiffvits_b200 through the drop-in API);
       reference       the UNMODIFIED reference (oracle/_ref snapshot: model3.Diffusion_Encoder incl. its per-step
                       PromptEncoder + sampler/*) on this host's CPU cores, on a bounded sample;
       gpu-reference   the same reference modules eager on cuda:0 (fp32, fp32+TF32, autocast-bf16): the "kernel bar".
The native run embeds the other two as `cpu_baseline` and `gpu_reference` (each measured in its own process, because the
reference and the drop-in packages share module names).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, T, Tp, sampler, steps, description)
    "c1": (1, 200, 282, "dpmsolver++", 50, "C1: 1 utterance x 200 frames, Tp=282, 50-step DPM-Solver++ 2M"),
    "c2": (32, 1000, 282, "dpmsolver++", 50, "C2: batch 32 x 1000 frames, Tp=282, 50-step DPM-Solver++ 2M"),
    "c3": (32, 1000, 282, "unipc", 10, "C3: batch 32 x 1000 frames, Tp=282, 10-step UniPC bh2 with prompt cross-attention"),
    "c4": (256, 1000, 282, "dpmsolver++", 50, "C4: GLOBAL batch 256 x 1000 frames batch-sharded over the GPUs, 50-step DPM-Solver++ 2M"),
    "c5": (8, 4000, 282, "unipc", 25, "C5: long-form batch 8 x 4000 frames, Tp=282, 25-step UniPC bh2"),
}
KIND_NAMES = {0: "gemm_tc", 1: "gn_stats", 2: "gn_apply", 3: "layernorm", 4: "attention", 5: "chain_tc"}


def denoiser_flops(T: int, Tp: int) -> float:
    """Algorithmic FLOPs of one reference denoiser call per utterance."""
    return 31.705e6 * T + 2.35e6 * Tp + 4352.0 * T * T + 7296.0 * T * Tp


def load_traffic() -> dict:
    """DRAM bytes per launch per kernel class from the committed ncu capture (profiles/traffic.json, written by
    tools/traffic_from_launches.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over one call)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def load_peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        d["_source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "_source": "fallback (B200_PROFILING.md)"}


def batch_per_gpu(args, world: int) -> int:
    Bg = WORKLOADS[args.workload][0]
    if args.scaling == "strong":
        gb = args.global_batch or Bg
        if gb % world:
            raise SystemExit(f"--global-batch {gb} is not divisible by {world} GPUs")
        return gb // world
    return args.global_batch or Bg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                for n, v in zip(names, parts[3:7]):
                    if v == "Active":
                        reasons.add(n)
        except Exception:
            pass
        finally:
            if self.path and os.path.exists(self.path):
                os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arms
class ReferenceJob:
    """The reference's own modules for the path, imported from the oracle/_ref snapshot (unmodified files):
    model3.Diffusion_Encoder (UNet1DConditionModel + PromptEncoder, evaluated once per denoiser call exactly as
    model3.py:902-914 does) driven by sampler/dpm_solver.py or sampler/uni_pc.py the way model3.sample wires them
    (model3.py:1143-1158 / 1177-1192). One instance per process (the reference's module names collide with the drop-in's)."""

    def __init__(self, device: str, threads: int):
        import torch
        from oracle import ref_snapshot
        ref_snapshot.activate("reference")
        self.torch = torch
        torch.set_num_threads(threads)
        self.model3 = ref_snapshot.import_model3()
        import sampler.dpm_solver as dpm
        import sampler.uni_pc as uni
        self.dpm, self.uni = dpm, uni
        cfg = ref_snapshot.load_config()
        torch.manual_seed(1234)
        self.diff = self.model3.Diffusion_Encoder(**cfg["diffusion_encoder"]).eval().to(device)
        self.device = device
        from oracle import inputs
        self.inputs = inputs
        self.betas = inputs.betas_linear().to(device)

    def make(self, B: int, T: int, Tp: int):
        x_T, cond, _ = self.inputs.make_inputs(B, T, Tp)
        prompt = self.inputs.make_prompt(B, Tp)
        dev = self.device
        lens = (self.torch.full((B,), T, device=dev), self.torch.full((B,), Tp, device=dev))
        return x_T.to(dev), (cond.to(dev), prompt.to(dev), lens[0], lens[1])

    def loop(self, x_T, data, sampler: str, steps_run: int):
        """`steps_run`-step loop; returns (latent, denoiser calls made)."""
        torch = self.torch
        calls = [0]

        def sample_fun(x, t, data=None):                     # model3.py:1113-1119 without the unused pred_noise
            calls[0] += 1
            return self.diff(x, data, t)

        with torch.no_grad():
            if sampler == "dpmsolver++":
                ns = self.dpm.NoiseScheduleVP("discrete", betas=self.betas)
                fn = self.dpm.model_wrapper(sample_fun, ns, model_type="x_start", model_kwargs={"data": data})
                lat = self.dpm.DPM_Solver(fn, ns, algorithm_type="dpmsolver++").sample(
                    x_T.clone(), steps=steps_run, order=2, skip_type="time_uniform", method="multistep")
            elif x_T.shape[0] == 1:
                ns = self.uni.NoiseScheduleVP("discrete", betas=self.betas)
                fn = self.uni.model_wrapper(sample_fun, ns, model_type="x_start", model_kwargs={"data": data})
                lat = self.uni.UniPC(fn, ns, variant="bh2").sample(
                    x_T.clone(), steps=steps_run, order=2, skip_type="time_uniform", method="multistep")
            else:
                # the reference's UniPC x_start wrapper broadcasts wrongly for B > 1 (uni_pc.py:189-191): the
                # batched loop is the oracle's restatement of UniPC bh2 around the reference's Diffusion_Encoder
                from oracle import sampler_oracle
                lat = sampler_oracle.unipc_bh2(lambda x, t_in: sample_fun(x, t_in, data=data), self.betas.cpu(),
                                               x_T.clone(), steps_run)
        return lat, calls[0]


def ref_sample_run(job: ReferenceJob, B, T, Tp, sampler, steps_total, steps_run):
    """Returns (seconds, frames/s extrapolated to the full `steps_total`-step loop over the sampled utterances)."""
    x_T, data = job.make(B, T, Tp)
    if job.device != "cpu":
        job.torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, calls = job.loop(x_T, data, sampler, steps_run)
    if job.device != "cpu":
        job.torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return dt, (B * T) / (dt / max(1, calls) * steps_total)


def _quiet_stdout() -> int:
    """stdout carries exactly ONE line (the JSON): what the reference / libraries print there goes to stderr. Returns the
    saved descriptor of the real stdout."""
    sys.stdout.flush()
    fd = os.dup(1)
    os.dup2(2, 1)
    return fd


def _emit(fd: int, line: dict) -> None:
    sys.stdout.flush()
    os.write(fd, (json.dumps(line) + "\n").encode())
    os.close(fd)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, timed on this host's cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    out_fd = _quiet_stdout()
    _, T, Tp, sampler, nfe, desc = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    Bg = batch_per_gpu(args, max(1, world))
    threads = os.cpu_count() or 1
    job = ReferenceJob("cpu", threads)
    Bs = min(Bg, 8)
    ref_sample_run(job, Bs, T, Tp, sampler, nfe, 2)           # first touch of the CPU libraries / allocator
    _, fps1 = ref_sample_run(job, Bs, T, Tp, sampler, nfe, 2)
    per_step = (Bs * T) / (fps1 * nfe)                        # seconds per solver step (denoiser call + update)
    steps_run = max(2, min(nfe, int(args.cpu_target / max(per_step, 1e-3))))
    for _ in range(args.warmup):
        ref_sample_run(job, Bs, T, Tp, sampler, nfe, 2)
    secs, fps = [], []
    for _ in range(args.steps):
        dt, f = ref_sample_run(job, Bs, T, Tp, sampler, nfe, steps_run)
        secs.append(dt)
        fps.append(f)
    value = sum(fps) / len(fps)
    sample = (f"{Bs} of {Bg} utterances x {T} frames, {steps_run}-step loop of the {nfe}-step {sampler} loop "
              f"({sum(secs) / len(secs):.1f} s of CPU work per timed step), extrapolated linearly in NFE and batch; the "
              f"reference's own model3.Diffusion_Encoder (UNet + per-step PromptEncoder) + sampler/* from the oracle/_ref "
              f"snapshot, FP32, eval, no_grad, all host threads")
    line = {
        "impl": "reference", "metric": "denoised latent frames/sec", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(secs) / len(secs), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "B_per_gpu": Bg, "T": T, "Tp": Tp, "sampler": sampler, "nfe": nfe},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(out_fd, line)


def run_gpu_reference(args):
    """--impl gpu-reference: the reference modules eager on cuda:0 (SURVEY 8d "kernel bar"): whole loop and one denoiser
    call in FP32 (TF32 off), FP32 + TF32, and under torch.autocast(bfloat16); one UNet call also CUDA-graph captured."""
    import torch
    out_fd = _quiet_stdout()
    _, T, Tp, sampler, nfe, desc = WORKLOADS[args.workload]
    Bg = batch_per_gpu(args, 1)
    job = ReferenceJob("cuda:0", os.cpu_count() or 1)
    x_T, data = job.make(Bg, T, Tp)
    res = {"device": torch.cuda.get_device_name(0), "workload": desc, "B": Bg,
           "what": "reference model3.Diffusion_Encoder (UNet + per-step PromptEncoder) + reference sampler, eager PyTorch "
                   "(cuDNN conv / cuBLAS / SDPA), same shapes and seeds as the native arm"}
    t_call = torch.full((Bg,), 500.0, device="cuda:0")

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    for mode in ("fp32", "fp32_tf32", "autocast_bf16"):
        tf32 = mode == "fp32_tf32"
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if mode == "autocast_bf16" else (lambda: torch.autocast("cuda", enabled=False))
        ent = {}

        def one_call():
            with torch.no_grad(), ctx():
                return job.diff(x_T, data, t_call)

        def one_loop():
            with ctx():
                return job.loop(x_T, data, sampler, nfe)

        try:
            ent["ms_per_denoiser_call"] = round(timed(one_call, 3), 3)
            ms_loop = timed(one_loop, 1)
            ent["ms_per_loop"] = round(ms_loop, 2)
            ent["frames_per_s"] = round(Bg * T / (ms_loop * 1e-3), 1)
            # one UNet call (prompt encoding outside) captured into a CUDA graph: the reference without launch overhead
            with torch.no_grad(), ctx():
                cond, prompt, _cl, pl = data
                mask = (torch.arange(Tp, device="cuda:0")[None, :] < pl[:, None])
                ehs = (job.diff.prompt_encoder(prompt, pl) * mask.unsqueeze(1).to(prompt.dtype)).transpose(1, 2).contiguous()
                xin = torch.cat([x_T, cond], 1)
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(2):
                        job.diff.unet(xin, t_call, ehs, encoder_attention_mask=mask)
                torch.cuda.current_stream().wait_stream(s)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    job.diff.unet(xin, t_call, ehs, encoder_attention_mask=mask)
            ent["ms_per_unet_call_cuda_graph"] = round(timed(gr.replay, 5), 3)
        except Exception as e:  # noqa: BLE001  (report, do not hide: the leg is a measurement of foreign code)
            ent["error"] = f"{type(e).__name__}: {e}"[:300]
        res[mode] = ent
    _emit(out_fd, {"impl": "gpu-reference", "gpu_reference": res})


def run_leg(impl: str, args, extra=()) -> dict:
    """Run another arm of this script in its own process and return its JSON line."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", impl, "--workload", args.workload, "--scaling", args.scaling,
           "--steps", "1", "--warmup", "0", *extra]
    if args.global_batch:
        cmd += ["--global-batch", str(args.global_batch)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (r.stderr or "no output")[-400:]}
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}


# ------------------------------------------------------------------------------------------------ native arm
def op_level(tag: str, n_levels: int = 4) -> int:
    """Resolution level (0 = T rows ... 3 = T/8 rows) of an op from its reference module path."""
    if tag.startswith("down_blocks."):
        i = int(tag.split(".")[1])
        return min(i + 1, n_levels - 1) if ".downsamplers." in tag else i
    if tag.startswith("up_blocks."):
        i = int(tag.split(".")[1])
        lvl = n_levels - 1 - i
        return lvl - 1 if ".upsamplers." in tag else lvl
    if tag.startswith("mid_block"):
        return n_levels - 1
    return 0


def run_native(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE line (rank 0's JSON): everything libraries print there (NCCL's version banner goes to
    # fd 1 from C) is sent to stderr by swapping the descriptors for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.precision != "bf16":
        os.environ["DV_PRECISION"] = args.precision
    import diffvits_b200
    from diffvits_b200 import synthetic as inputs       # the native arm never imports oracle/
    from diffvits_b200.synthetic import UNET_KW
    diffvits_b200.install_dropin()
    import sampler.dpm_solver as dpm
    import sampler.uni_pc as uni
    from unet1d.unet_1d_condition import UNet1DConditionModel

    _, T, Tp, sampler, nfe, desc = WORKLOADS[args.workload]
    Bg = batch_per_gpu(args, world)
    torch.manual_seed(1234)
    unet = UNet1DConditionModel(**UNET_KW).eval().to(dev)
    # noise / conditioning are drawn for the GLOBAL batch with one seed, then sharded (bit-identical to 1 GPU)
    from diffvits_b200.dist import gather_latents, init_native_comm, shard_inputs
    gather_how = "single GPU: no collective"
    if world > 1:
        gather_how = ("ncclAllGather issued in-stream through the C ABI (dv_comm_all_gather, library-owned communicator)"
                      if init_native_comm(dev) and not os.environ.get("DV_TORCH_GATHER")
                      else "torch.distributed.all_gather_into_tensor (NCCL)")
        if os.environ.get("DV_TORCH_GATHER"):
            from diffvits_b200.dist import close_native_comm
            close_native_comm()
    x_h, cond_h, _unused = (t.pin_memory() for t in shard_inputs(inputs.make_inputs(Bg * world, T, Tp), rank, world))
    # the speech prompt (mel, model3.py:906) and its encoder: encoded ONCE per job by the native PromptEncoder plan
    # (the reference re-encodes it in every denoiser call); random-init parameters in the reference's state_dict layout
    from diffvits_b200.prompt import PromptEncoderParams, encode_prompt
    prompt_h = shard_inputs((inputs.make_prompt(Bg * world, Tp),), rank, world)[0].pin_memory()
    penc = PromptEncoderParams().eval().to(dev)
    plen_d = torch.full((Bg,), Tp, dtype=torch.int32, device=dev)
    betas = inputs.betas_linear()
    x_d, cond_d = x_h.to(dev), cond_h.to(dev)
    with torch.no_grad():
        ehs_d = encode_prompt(penc, prompt_h.to(dev), plen_d)
    assert ehs_d is not None
    native = diffvits_b200.NativeDenoiser(unet, cond_d, ehs_d, lanes=args.lanes)
    eng = native.engine

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, iters):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / iters

    # ---- device-resident loop: inputs already in HBM, whole loop = one CUDA-graph replay
    def step_resident():
        lat = diffvits_b200.sample_native(unet, x_d, cond_d, ehs_d, sampler=sampler, steps=nfe, betas=betas,
                                          native=native)
        return gather_latents(lat, Bg * world)

    for _ in range(max(3, args.warmup)):
        step_resident()
    with ClockSampler(local) as cs:
        ms_res = timed(step_resident, args.steps)
    clocks = cs.summary()
    launches_per_step = eng.last_sample_launches + (1 if world > 1 else 0)

    # ---- end to end through the reference-facing API (drop-in model_wrapper + solver), host buffers in and out
    if sampler == "dpmsolver++":
        ns = dpm.NoiseScheduleVP("discrete", betas=betas)
        solver = dpm.DPM_Solver(dpm.model_wrapper(native, ns, model_type="x_start", model_kwargs={}), ns,
                                algorithm_type="dpmsolver++")
    else:
        ns = uni.NoiseScheduleVP("discrete", betas=betas)
        solver = uni.UniPC(uni.model_wrapper(native, ns, model_type="x_start", model_kwargs={}), ns, variant="bh2")
    out_h = torch.empty(Bg, 100, T).pin_memory()

    def step_e2e():
        xd = x_h.to(dev, non_blocking=True)
        cd = cond_h.to(dev, non_blocking=True)
        pd = prompt_h.to(dev, non_blocking=True)
        with torch.no_grad():
            ed = encode_prompt(penc, pd, plen_d)                  # PromptEncoder, once per job (native plan)
        native.rebind(cd, ed)                                     # hoisted prompt work is part of the job
        lat = solver.sample(xd, steps=nfe, order=2, skip_type="time_uniform", method="multistep")
        full = gather_latents(lat, Bg * world)
        out_h.copy_(full[rank * Bg:(rank + 1) * Bg], non_blocking=True)
        torch.cuda.current_stream().synchronize()                 # the caller holds the latent on the host
        assert solver.last_run_native

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    h2d = x_h.numel() * 4 + cond_h.numel() * 4 + prompt_h.numel() * 4
    d2h = out_h.numel() * 4

    frames = Bg * world * T
    value = frames / (ms_res * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        peaks = load_peaks()
        peak_tf, peak_hbm = peaks["bf16_tflops"], peaks["hbm_gbs"]
        # ---- per-launch kernel durations: one denoiser call replayed as a CUDA graph with an event-record node between
        # consecutive launches (dv_plan_profile_forward), CUDA events on the launching stream, mean of `reps` replays
        ops = eng.op_table()
        t_probe = torch.tensor([500.0], device=dev)
        eng.profile_forward(x_d, t_probe)                         # warm
        acc = [0.0] * (len(ops) + 1)
        reps = 3
        for _ in range(reps):
            for i, v in enumerate(eng.profile_forward(x_d, t_probe)):
                acc[i] += v / reps
        by_kind, by_level = {}, {}
        for o, ms in zip(ops, acc[:-1]):
            k = by_kind.setdefault(KIND_NAMES[o["kind"]], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0, "roof_ms": 0.0})
            t_tensor = o["flops"] / (peak_tf * 1e12) * 1e3          # ms at the tensor roof
            t_hbm = o["bytes"] / (peak_hbm * 1e9) * 1e3             # ms at the HBM roof (algorithmic bytes)
            k["ms"] += ms
            k["flops"] += o["flops"]
            k["bytes"] += o["bytes"]
            k["launches"] += 1
            k["roof_ms"] += max(t_tensor, t_hbm)
            if o["kind"] == 0:
                lv = by_level.setdefault((op_level(o["tag"]), "tensor" if t_tensor >= t_hbm else "hbm"),
                                         {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0, "roof_ms": 0.0})
                lv["ms"] += ms
                lv["flops"] += o["flops"]
                lv["bytes"] += o["bytes"]
                lv["launches"] += 1
                lv["roof_ms"] += max(t_tensor, t_hbm)
        call_ms_eager = sum(acc)
        n_fused_gn = sum(1 for o in ops if "+gn:" in o["tag"])
        g = by_kind["gemm_tc"]
        gemm_tflops = g["flops"] / (g["ms"] * 1e-3) / 1e12
        kernels = {}
        for name, k in by_kind.items():
            ent = {"launches_per_call": k["launches"], "ms_per_call": round(k["ms"], 4),
                   "share_of_call": round(k["ms"] / call_ms_eager, 4), "us_per_launch": round(1e3 * k["ms"] / k["launches"], 2)}
            if k["flops"] > 0:
                ent["tflops"] = round(k["flops"] / (k["ms"] * 1e-3) / 1e12, 2)
                ent["frac_of_bf16_peak"] = round(ent["tflops"] / peak_tf, 4)
                ent["frac_of_binding_roof"] = round(k["roof_ms"] / k["ms"], 4)
            if name in ("gn_stats", "gn_apply", "layernorm"):
                ent["gbs"] = round(k["bytes"] / (k["ms"] * 1e-3) / 1e9, 1)
                ent["frac_of_hbm_peak"] = round(ent["gbs"] / peak_hbm, 4)
            kernels[name] = ent
        kernels["prologue(layout+time_mlp)"] = {"ms_per_call": round(acc[-1], 4)}
        gemm_levels = {}
        for (lvl, bound), lv in sorted(by_level.items()):
            gemm_levels[f"L{lvl}/{bound}-bound"] = {
                "launches": lv["launches"], "ms_per_call": round(lv["ms"], 4),
                "tflops": round(lv["flops"] / (lv["ms"] * 1e-3) / 1e12, 1),
                "algorithmic_gbs": round(lv["bytes"] / (lv["ms"] * 1e-3) / 1e9, 1),
                "frac_of_binding_roof": round(lv["roof_ms"] / lv["ms"], 4)}

        # ---- sampler update kernel (north_star item 3), timed alone over buffers that do not fit L2 together
        from diffvits_b200 import _lib
        lib = _lib.load()
        n_el = Bg * 100 * T
        nset = max(2, int(160e6 // (4 * n_el * 4)) + 1)
        bufs = [[torch.randn(n_el, device=dev) for _ in range(3)] for _ in range(nset)]
        st = torch.cuda.current_stream().cuda_stream

        def upd():
            for x, m1, m0 in bufs:
                lib.dv_dpmpp_update(x.data_ptr(), x.data_ptr(), m1.data_ptr(), m0.data_ptr(), 0.9, 0.05, 0.05, 0, 1.0, 1.0,
                                    n_el, st)
        upd()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            upd()
        e1.record()
        torch.cuda.synchronize(dev)
        upd_us = e0.elapsed_time(e1) / (5 * nset) * 1e3
        upd_gbs = 4.0 * n_el * 4 / (upd_us * 1e-6) / 1e9
        kernels["sampler_update(dpmpp_2m)"] = {
            "us_per_launch": round(upd_us, 2), "algorithmic_bytes": 16 * n_el, "gbs": round(upd_gbs, 1),
            "frac_of_hbm_peak": round(upd_gbs / peak_hbm, 4),
            "how": f"{nset} buffer sets ({nset * 3 * 4 * n_el / 1e6:.0f} MB > L2) launched back to back, CUDA events"}
        del bufs

        # ---- optional (debug library only): the same classes INSIDE the timed loop by leaving one class out
        in_loop = {}
        if hasattr(lib, "dv_debug_skip_kinds"):
            def step_local():
                diffvits_b200.sample_native(unet, x_d, cond_d, ehs_d, sampler=sampler, steps=nfe, betas=betas, native=native)

            def timed_local(iters=2):
                step_local()
                step_local()
                torch.cuda.synchronize(dev)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(iters):
                    step_local()
                a1.record()
                torch.cuda.synchronize(dev)
                return a0.elapsed_time(a1) / iters

            ms_all = timed_local()
            for name, kind in (("gemm_tc", 0), ("gn_apply", 2), ("layernorm", 3), ("attention", 4), ("chain_tc", 5)):
                if name in kernels:
                    lib.dv_debug_skip_kinds(1 << kind)
                    in_loop[name] = max(ms_all - timed_local(), 1e-6) / nfe
            lib.dv_debug_skip_kinds(0)
            step_local()
            torch.cuda.synchronize(dev)
            for name, ms in in_loop.items():
                kernels[name]["in_loop_ms_per_call_by_subtraction"] = round(ms, 4)

        total_flops = denoiser_flops(T, Tp) * Bg * nfe           # per GPU per step
        step_tflops = total_flops / (ms_res * 1e-3) / 1e12
        traffic = load_traffic()
        for name, ent in kernels.items():
            if name in traffic.get("per_launch_dram_bytes", {}):
                ent["ncu_dram_bytes_per_launch"] = traffic["per_launch_dram_bytes"][name]
        roofline = {
            "kernel": "gemm_tc_kernel (tcgen05 tap-GEMM: conv1d k3 / 1x1 / linear)",
            "bound": "tensor", "achieved": round(gemm_tflops, 2), "peak": peak_tf, "unit": "TFLOP/s",
            "frac": round(gemm_tflops / peak_tf, 4),
            "traffic": traffic.get("per_launch_dram_bytes", {}).get("gemm_tc"),
            "traffic_source": traffic.get("source"),
            "algorithmic_bytes_per_launch": round(g["bytes"] / g["launches"]),
            "algorithmic_flops_per_launch": round(g["flops"] / g["launches"]),
            "us_per_launch": round(1e3 * g["ms"] / g["launches"], 2),
            "peak_source": peaks["_source"] + ", burst figure (kernel timed launch by launch)",
            "binding_roof": {
                "frac": round(g["roof_ms"] / g["ms"], 4),
                "what": "sum over launches of max(FLOPs / bf16 peak, algorithmic bytes / HBM peak) divided by the summed "
                        "kernel durations: level-0 shapes (128 channels) are HBM-bound, deeper levels tensor-bound",
                "by_level": gemm_levels},
            "how": f"achieved = algorithmic FLOPs of the {g['launches']} gemm_tc launches of one denoiser call / the SUM OF "
                   f"THEIR KERNEL DURATIONS: the call replayed as a CUDA graph with an event-record node between launches, "
                   f"CUDA events on the launching stream, mean of {reps} replays (each launch pays its own fill and drain, "
                   f"no PDL overlap with its neighbours)" +
                   (f"; {n_fused_gn} of these launches also apply the GroupNorm (+SiLU, +scale/shift) of their own output in "
                    f"their epilogue (per-item arrival counter): that tail is inside the durations while no FLOPs are counted "
                    f"for it (DV_GN_OUT=0 runs it as {n_fused_gn} stand-alone gn_apply launches instead)" if n_fused_gn else ""),
            "fused_groupnorm_launches": n_fused_gn,
            "whole_step": {"algorithmic_tflop_per_step_per_gpu": round(total_flops / 1e12, 2),
                           "achieved_tflops": round(step_tflops, 2),
                           "frac_of_burst_peak": round(step_tflops / peak_tf, 4),
                           "frac_of_sustained_peak": round(step_tflops / peaks["bf16_tflops_sustained"], 4)},
            "kernels": kernels,
        }
        line = {
            "metric": "denoised latent frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_res, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "fp16", "fp16s": "fp16", "fp32x3": "bf16x3"}[args.precision],
            "data": "synthetic",
            "rtf": (ms_res * 1e-3) / (frames * 256 / 24000.0),
            "config": {"workload": desc + {"bf16": ", BF16 operands / FP32 accumulate",
                                           "fp16": ", FP16 operands / FP32 accumulate",
                                           "fp16s": ", FP16 operands / FP32 accumulate, hi+lo residual stream",
                                           "fp32x3": ", hi+lo BF16 operand split (3 MMAs), FP32-class"}[args.precision]
                                   + ", FP32 sampler state",
                       "B_per_gpu": Bg, "global_batch": Bg * world, "T": T, "Tp": Tp, "sampler": sampler, "nfe": nfe,
                       "order": 2, "weights": "random init seed 1234 (64.7 M params)", "inputs": "N(0,1) seed 0",
                       "l2": "no flush: per-step working set (workspace %.2f GB + weights 0.15 GB) >> 126 MB L2" %
                             (eng.ws_bytes / 1e9),
                       "gather": gather_how,
                       "graph": "whole loop = one CUDA graph (%d kernel nodes, %d concurrent lanes of %s utterances)" %
                                (eng.last_sample_launches, eng.lanes, "/".join(map(str, eng.sizes)))},
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "path": "pinned host x_T / cond / prompt mel -> H2D -> native PromptEncoder (once per job) -> "
                            "NativeDenoiser.rebind (prompt K/V, text-time emb) -> drop-in model_wrapper + solver.sample "
                            "(native graph) -> D2H latent"},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "roofline": roofline,
        }
    # ---- the other precision modes on the same workload (device-resident loop; parity of each mode: tests/test_gpu_precision.py)
    if rank == 0 and world == 1 and not args.no_modes and args.precision == "bf16":
        import copy
        modes = {"bf16": {"ms_per_step": round(ms_res, 2), "frames_per_s": round(value, 1),
                          "parity_vs_fp32_reference": "per call 8.3e-3 rel-L2 / 1.4e-2 max-abs (the reference under autocast-bf16: 1.05e-2)"}}
        notes = {"fp16": "FP16 operands in place of the BF16 ones (same launches, chains and fused norms): per call 1.0e-3 rel-L2 / "
                         "1.6e-3 max-abs, 10-step UniPC latent 1.1e-3, 50-step DPM-Solver++ latent 2.6e-3",
                 "fp16s": "FP16 operands + hi/lo residual stream, one MMA per product: per call 8.1e-4 rel-L2 / 1.3e-3 max-abs, "
                          "10-step UniPC latent 9.0e-4, 50-step DPM-Solver++ latent 1.9e-3",
                 "fp32x3": "hi+lo BF16 split, three MMAs per product (GEMMs and attention): per call 1.5e-5 rel-L2, 50-step latent 3.8e-5"}
        for mode in ("fp16", "fp16s", "fp32x3"):
            u2 = copy.deepcopy(unet)
            u2.native_precision = mode
            n2 = diffvits_b200.NativeDenoiser(u2, cond_d, ehs_d)

            def step2():
                diffvits_b200.sample_native(u2, x_d, cond_d, ehs_d, sampler=sampler, steps=nfe, betas=betas, native=n2)

            for _ in range(3):
                step2()
            ms2 = timed(step2, 2)
            modes[mode] = {"ms_per_step": round(ms2, 2), "frames_per_s": round(frames / (ms2 * 1e-3), 1),
                           "parity_vs_fp32_reference": notes[mode]}
            n2.close()
            del n2, u2
        line["precision_modes"] = modes
    if world > 1:
        dist.barrier()
        from diffvits_b200.dist import close_native_comm
        close_native_comm()
        dist.destroy_process_group()
    if rank == 0:
        # free this process's GPU memory before the reference legs run in their own processes
        del native, eng, solver
        torch.cuda.empty_cache()
        if world == 1 and not args.no_cpu:
            ref = run_leg("reference", args, ["--cpu-target", "15"])
            line["cpu_baseline"] = ref.get("cpu_baseline", {"error": ref.get("error", "no cpu_baseline in the reference line")})
        else:
            line["cpu_baseline"] = None
        if world == 1 and not args.no_gpu_ref:
            gref = run_leg("gpu-reference", args)
            line["gpu_reference"] = gref.get("gpu_reference", {"error": gref.get("error", "?")})
            best = [v.get("ms_per_loop") for v in line["gpu_reference"].values() if isinstance(v, dict) and v.get("ms_per_loop")]
            if best:
                line["gpu_reference"]["native_speedup_vs_best_eager_loop"] = round(min(best) / ms_res, 2)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference", "gpu-reference"])
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak: the workload's batch per GPU (default); strong: a fixed global batch split over the GPUs "
                         "(default for --workload c4)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong: global batch (default: the workload's); weak: overrides the per-GPU batch")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp16s", "fp32x3"],
                    help="bf16: BF16 operands (headline); fp16: FP16 operands, same program; fp16s: FP16 operands + hi/lo "
                         "residual stream (1e-3 gate); fp32x3: hi+lo BF16 operand split, FP32-class accuracy")
    ap.add_argument("--cpu-target", type=float, default=10.0, help="reference arm: seconds of CPU work per timed step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-gpu-ref", action="store_true", help="skip the gpu_reference leg")
    ap.add_argument("--no-modes", action="store_true", help="skip the fp16s / fp32x3 timing legs")
    ap.add_argument("--lanes", type=int, default=None,
                    help="concurrent batch slices inside the loop graph (default: engine heuristic / DV_LANES)")
    args = ap.parse_args()
    if args.scaling is None:
        args.scaling = "strong" if args.workload == "c4" else "weak"
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "gpu-reference":
        run_gpu_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
