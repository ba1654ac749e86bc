#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 2 --launch-count 1 -o gpurun_out/prof_one -f python tools/ncu_one_gemm.py 0 > gpurun_out/ncu_one.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 2 --launch-count 1 -o gpurun_out/prof_one6 -f python tools/ncu_one_gemm.py 62 > gpurun_out/ncu_one6.log 2>&1; echo "rc=$?"
ls -la gpurun_out/prof_one*.ncu-rep
