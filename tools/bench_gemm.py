"""Micro-benchmark of the tcgen05 tap-GEMM through the C ABI: fixed vs per-tile cost, by shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as U

def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, generator=g, device="cuda")).to(torch.bfloat16)
print("shape(B,L,K,N,taps,mode,bn)  us  TFLOPS  tiles")
for (B, L, K, N, taps, mode, bn) in [
    (16, 1000, 128, 128, 1, 0, 128), (32, 1000, 128, 128, 1, 0, 128), (64, 1000, 128, 128, 1, 0, 128), (128, 1000, 128, 128, 1, 0, 128),
    (32, 1000, 128, 128, 3, 0, 128), (64, 1000, 128, 128, 3, 0, 128),
    (32, 1000, 128, 384, 1, 0, 128), (32, 1000, 128, 1024, 1, 2, 128), (32, 1000, 128, 1024, 1, 2, 256), (32, 1000, 128, 1024, 1, 0, 128), (32, 1000, 128, 1024, 1, 0, 256),
    (32, 1000, 512, 128, 1, 0, 128),
    (32, 500, 256, 256, 3, 0, 128), (32, 500, 256, 256, 3, 0, 256), (32, 500, 256, 2048, 1, 2, 256), (32, 500, 256, 2048, 1, 2, 128),
    (32, 250, 384, 384, 3, 0, 128), (32, 125, 512, 512, 3, 0, 128), (32, 125, 512, 512, 3, 0, 256),
    (128, 1000, 384, 512, 1, 0, 256), (128, 1000, 384, 512, 1, 0, 128), (64, 1000, 1024, 1024, 1, 0, 256),
]:
    a = rnd(B, L, K)
    w = rnd(N, K * taps)
    segs = [(0, t - (taps // 2), 0, K) for t in range(taps)]
    res = rnd(B, L, N) if mode == 0 else None
    bias = torch.randn(N, device="cuda")
    fn = lambda: U.gemm([a], segs, w, N, bias=bias, res=res, out_mode=mode, bn=bn)
    us = timeit(fn)
    flops = 2.0 * B * L * N * K * taps
    tiles = B * ((L + 127) // 128) * (N // bn)
    print(f"{(B, L, K, N, taps, mode, bn)}  {us:8.1f}  {flops / us / 1e6:7.1f}  {tiles}")
