#!/bin/bash
# attention kernel A/B: parity tests for every variant, then the per-shape microbench per variant ("DV_ATTN:DV_ATTN_POLY")
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k attention > gpurun_out/attn_tests.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/attn_tests.log
for v in ${@:-legacy:0 v2:0 v2:3}; do
  echo "== DV_ATTN=${v%%:*} DV_ATTN_POLY=${v##*:}"
  DV_ATTN=${v%%:*} DV_ATTN_POLY=${v##*:} timeout 300 python tools/microbench.py --only attn 2>&1 | grep attn
done
