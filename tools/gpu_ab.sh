#!/bin/bash
# generic A/B: kernel tests, then bench.py --no-cpu once per env setting given as arguments ("VAR=val VAR2=val" strings; "-" = defaults)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_chain.py -q -m gpu -x > gpurun_out/ab_tests.log 2>&1; echo "kernel tests rc=$?"; tail -n 3 gpurun_out/ab_tests.log
for cfg in "$@"; do
  [ "$cfg" = "-" ] && cfg=""
  echo "== bench [$cfg]"
  env $cfg timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],2), 'ms_per_call', round(d['ms_per_step']/50,4), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
