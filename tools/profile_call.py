"""Per-launch profile of ONE denoiser call (eager enqueue, CUDA events between launches) + a target for ncu:
    python tools/profile_call.py [--workload c2] [--out gpurun_out/ops_c2.csv] [--calls 2]
ncu usage (one GPU):  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... \
                          --nvtx --nvtx-include "dv_call/" python tools/profile_call.py --calls 1 --no-events
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--out", default=None)
    ap.add_argument("--calls", type=int, default=3)
    ap.add_argument("--no-events", action="store_true")
    a = ap.parse_args()
    import diffvits_b200
    from diffvits_b200 import synthetic as inputs
    from diffvits_b200.synthetic import UNET_KW
    diffvits_b200.install_dropin()
    from unet1d.unet_1d_condition import UNet1DConditionModel
    Bg, T, Tp, sampler, nfe, desc = bench.WORKLOADS[a.workload]
    torch.manual_seed(1234)
    unet = UNet1DConditionModel(**UNET_KW).eval().cuda()
    x, cond, ehs = (t.cuda() for t in inputs.make_inputs(Bg, T, Tp))
    native = diffvits_b200.NativeDenoiser(unet, cond, ehs)
    eng = native.engine
    t = torch.tensor([500.0], device="cuda")
    native(x, t)
    torch.cuda.synchronize()
    if a.no_events:
        torch.cuda.nvtx.range_push("dv_call")
        for _ in range(a.calls):
            native(x, t)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return
    ops = eng.op_table()
    acc = [0.0] * (len(ops) + 1)
    for _ in range(a.calls):
        for i, v in enumerate(eng.profile_forward(x, t)):
            acc[i] += v / a.calls
    rows = ["idx,tag,kind,ms,gflop,mbytes,tflops,gbs"]
    for i, (o, ms) in enumerate(zip(ops, acc)):
        rows.append(f"{i},{o['tag']},{bench.KIND_NAMES[o['kind']]},{ms:.5f},{o['flops'] / 1e9:.3f},{o['bytes'] / 1e6:.3f},"
                    f"{o['flops'] / (ms * 1e-3) / 1e12:.1f},{o['bytes'] / (ms * 1e-3) / 1e9:.0f}")
    rows.append(f"{len(ops)},prologue,misc,{acc[-1]:.5f},0,0,0,0")
    text = "\n".join(rows)
    if a.out:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        open(a.out, "w").write(text + "\n")
    print(f"# {desc}: graph-replayed call {sum(acc):.3f} ms over {len(ops)} launches")
    print(text)


if __name__ == "__main__":
    main()
