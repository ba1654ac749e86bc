#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
tail -n 30 gpurun_out/t_kernels.log
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -s > gpurun_out/t_unet.log 2>&1; echo "unet rc=$?"
grep -E "passed|failed|FAILED|Error|final latent|t=" gpurun_out/t_unet.log | head -30
timeout 300 python tools/profile_call.py --workload c2 --out gpurun_out/ops_c2.csv > gpurun_out/profile_c2.log 2>&1; echo "profile rc=$?"; head -n 1 gpurun_out/profile_c2.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d['roofline']['kernels'], indent=0))"
