#!/bin/bash
# library bar: every kernel class at the C2 shapes next to the PyTorch library op (cuDNN / cuBLASLt / SDPA), + one test
cd "$(dirname "$0")/.."
TAG=${1:-lb}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -q -x -p no:cacheprovider -k "groupnorm_in_producer" > gpurun_out/${TAG}_tests1.log 2>&1; echo "pytest gn rc=$?"; tail -2 gpurun_out/${TAG}_tests1.log
timeout 900 python tools/microbench.py --lib --taps3 > gpurun_out/${TAG}_library_bar_c2.txt 2>&1; echo "microbench rc=$?"; tail -50 gpurun_out/${TAG}_library_bar_c2.txt | cut -c1-220
