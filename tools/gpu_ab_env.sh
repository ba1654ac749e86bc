#!/bin/bash
# A/B of one environment switch on the C2 step, alternating runs on one box: usage gpu_ab_env.sh VAR valA valB [reps]
cd "$(dirname "$0")/.."
VAR=$1; A=$2; B=$3; REPS=${4:-2}
mkdir -p gpurun_out
for r in $(seq 1 $REPS); do
for v in "$A" "$B"; do
env $VAR=$v timeout 600 python bench.py --steps 3 --no-cpu --no-gpu-ref --no-modes > gpurun_out/ab_tmp.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/ab_tmp.json').read().strip().splitlines()[-1])
print('$VAR=$v step ms', round(d['ms_per_step'],2), 'gemm', d['roofline']['kernels']['gemm_tc']['ms_per_call'])
"
done
done
