"""Cost of a NEW shape on the drop-in path (tts_infer.py synthesises one utterance of a new length per call):
plan build + workspace + bind (NativeDenoiser), first sample (graph capture + instantiate + run), replay.
    python tools/first_call_latency.py [--B 1] [--steps 50]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    import diffvits_b200
    from diffvits_b200 import synthetic as inputs
    from diffvits_b200.synthetic import UNET_KW
    diffvits_b200.install_dropin()
    from unet1d.unet_1d_condition import UNet1DConditionModel
    torch.manual_seed(1234)
    unet = UNet1DConditionModel(**UNET_KW).eval().cuda()
    betas = inputs.betas_linear()

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    print("T, Tp: denoiser ctor ms | first sample ms | second sample ms | eager (no graph) sample ms")
    for T, Tp in ((400, 282), (403, 282), (417, 300), (1000, 282), (403, 282)):
        x, c, e = (v.cuda() for v in inputs.make_inputs(a.B, T, Tp, seed=T))
        t0 = sync()
        native = diffvits_b200.NativeDenoiser(unet, c, e)
        t1 = sync()
        diffvits_b200.sample_native(unet, x, c, e, sampler="dpmsolver++", steps=a.steps, betas=betas, native=native)
        t2 = sync()
        diffvits_b200.sample_native(unet, x, c, e, sampler="dpmsolver++", steps=a.steps, betas=betas, native=native)
        t3 = sync()
        try:
            diffvits_b200.sample_native(unet, x, c, e, sampler="dpmsolver++", steps=a.steps, betas=betas, native=native,
                                        graph=False)
            t4 = sync()
            eager = f"{(t4 - t3) * 1e3:8.1f}"
        except TypeError:
            eager = "   n/a"
        print(f"{T:5d} {Tp:4d}: {(t1 - t0) * 1e3:8.1f} | {(t2 - t1) * 1e3:8.1f} | {(t3 - t2) * 1e3:8.1f} | {eager}")


if __name__ == "__main__":
    main()
