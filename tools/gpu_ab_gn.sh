#!/bin/bash
# A/B of the GEMM pipeline shapes on the C2 step: one-ring kernel (default), two-ring kernel with shared conv taps (DV_GEMM2=1),
# two-ring kernel with GroupNorm fused into the operand path (DV_FUSE_GN=1)
run() { echo "== $*"; env "$@" python bench.py --steps 3 --no-cpu --no-gpu-ref 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernels']; print('step ms', round(d['ms_per_step'],2), {n:(v.get('launches_per_call'), v.get('ms_per_call')) for n,v in k.items() if 'ms_per_call' in v and n[:3] in ('gem','gn_','lay')})"; }
run DV_GEMM2=0
run DV_GEMM2=1
run DV_FUSE_GN=1
