#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dv_call/" -k regex:gemm_tc --launch-skip 1 --launch-count 10 -o gpurun_out/prof_gemm -f python tools/profile_call.py --workload c2 --calls 1 --no-events > gpurun_out/ncu_gemm.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_gemm.log; ls -la gpurun_out/prof_gemm.ncu-rep
