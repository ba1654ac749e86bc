#!/bin/bash
# usage: gpu_quick.sh tag "<microbench args>" [unet]
cd "$(dirname "$0")/.."
TAG=${1:-q}
mkdir -p gpurun_out
if [ "$3" = "unet" ]; then
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -s > gpurun_out/t_unet_$TAG.log 2>&1; echo "unet rc=$?"
grep -E "passed|failed|FAILED|Error|final latent|t=" gpurun_out/t_unet_$TAG.log | head -20
fi
timeout 600 python tools/microbench.py $2 > gpurun_out/micro_$TAG.log 2>&1; echo "micro rc=$?"
cat gpurun_out/micro_$TAG.log
