# Needs the debug library: DV_DEBUG=1 python -m diffvits_b200.build ; run with DV_DEBUG_LIB=1 (include/diffvits_b200_debug.h)
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_util as U
from diffvits_b200 import _lib
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*s): return (torch.randn(*s, generator=g, device="cuda")).to(torch.bfloat16)
B, L, K, N = 512, 1000, 128, 128
a = rnd(B, L, K); w = rnd(N, K); bias = torch.randn(N, device="cuda"); res = rnd(B, L, N)
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib.dv_gemm_test_flags(flags)
for _ in range(3):
    U.gemm([a], [(0, 0, 0, K)], w, N, bias=bias, res=res, out_mode=0, bn=128)
torch.cuda.synchronize()
