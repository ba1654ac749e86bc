"""Which kernel does torch's F.scaled_dot_product_attention run for the C2 level-0 self-attention shape (the library bar of
tools/microbench.py --lib)? Prints the CUDA kernel names seen by the profiler."""
import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

for (B, H, L, D) in ((32, 8, 1000, 16), (32, 8, 500, 32), (32, 8, 125, 64)):
    q, k, v = (torch.randn(B, H, L, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    for _ in range(3):
        F.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        F.scaled_dot_product_attention(q, k, v)
        torch.cuda.synchronize()
    for e in prof.key_averages():
        if e.device_time_total > 0:
            print(f"B={B} H={H} L={L} D={D}: {e.key[:150]}  {e.device_time_total:.1f} us")
