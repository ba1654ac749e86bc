#!/bin/bash
# first GPU contact: kernel parity, then the end-to-end parity tests; logs under gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,driver_version --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm_linear" -x > gpurun_out/t1_gemm.log 2>&1; echo "gemm_linear rc=$?"
tail -n 25 gpurun_out/t1_gemm.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/t2_kernels.log 2>&1; echo "kernels rc=$?"
tail -n 40 gpurun_out/t2_kernels.log
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -s > gpurun_out/t3_unet.log 2>&1; echo "unet rc=$?"
tail -n 60 gpurun_out/t3_unet.log
