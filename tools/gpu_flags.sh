#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in 1 6 2; do
timeout 300 python tools/microbench.py --only gemm --flags $f > gpurun_out/micro_flags$f.log 2>&1; echo "flags $f rc=$?"
grep -E "^L0|^L1" gpurun_out/micro_flags$f.log
done
