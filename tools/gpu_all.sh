#!/bin/bash
# One gpurun call: parity tests, micro-benchmarks, per-op profile, bench line.  usage: gpu_all.sh [tag] [steps]
cd "$(dirname "$0")/.."
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/gpu_$TAG.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/t_kernels_$TAG.log 2>&1; echo "kernels rc=$?"
tail -n 12 gpurun_out/t_kernels_$TAG.log
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -s > gpurun_out/t_unet_$TAG.log 2>&1; echo "unet rc=$?"
grep -E "passed|failed|FAILED|Error|final latent|t=" gpurun_out/t_unet_$TAG.log | head -20
timeout 600 python tools/microbench.py > gpurun_out/micro_$TAG.log 2>&1; echo "micro rc=$?"
cat gpurun_out/micro_$TAG.log
timeout 300 python tools/profile_call.py --workload c2 --out gpurun_out/ops_c2_$TAG.csv > gpurun_out/profile_c2_$TAG.log 2>&1; echo "profile rc=$?"; head -n 1 gpurun_out/profile_c2_$TAG.log
timeout 900 python bench.py --steps ${2:-3} --warmup 3 ${3:---no-cpu} > gpurun_out/bench_c2_$TAG.json 2> gpurun_out/bench_c2_$TAG.err; echo "bench rc=$?"
tail -n 3 gpurun_out/bench_c2_$TAG.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_$TAG.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step']); print(json.dumps(d['roofline']['kernels'])); print(d.get('cpu_baseline'))"
