#!/bin/bash
# One gpurun call (1 GPU): ncu launch list of one denoiser call + `--set full` captures of representative launches.
# usage: gpu_profile.sh [tag]      outputs: gpurun_out/launches_<tag>.csv, gpurun_out/prof_*_<tag>.ncu-rep
cd "$(dirname "$0")/.."
TAG=${1:-r1}
mkdir -p gpurun_out
# every launch of one eager denoiser call (C2 shapes), device time per launch (cold cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  --nvtx --nvtx-include "dv_call/" python tools/profile_call.py --workload c2 --calls 1 --no-events \
  > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?"
# full sections: first resnet + first transformer block at level 0 (conv, gn, ln, qkv, attention, geglu ...)
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dv_call/" \
  --launch-skip 3 --launch-count 22 -o gpurun_out/prof_l0_$TAG -f \
  python tools/profile_call.py --workload c2 --calls 1 --no-events > gpurun_out/ncu_l0_$TAG.log 2>&1; echo "ncu l0 rc=$?"
# sampler update kernel (HBM-bound), C2 state size
timeout 600 ncu --set full --clock-control none --import-source on -k regex:axpbypcz --launch-skip 25 --launch-count 2 \
  -o gpurun_out/prof_sampler_$TAG -f python tools/microbench.py --only samp > gpurun_out/ncu_sampler_$TAG.log 2>&1; echo "ncu sampler rc=$?"
ls -la gpurun_out/*.ncu-rep
