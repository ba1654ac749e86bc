#!/bin/bash
# several values of one environment switch on the C2 step, one box: usage gpu_ab_multi.sh VAR "v1 v2 v3 ..."
cd "$(dirname "$0")/.."
VAR=$1; VALS=$2
mkdir -p gpurun_out
for v in $VALS; do
env $VAR=$v timeout 600 python bench.py --steps 3 --no-cpu --no-gpu-ref --no-modes > gpurun_out/ab_tmp.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/ab_tmp.json').read().strip().splitlines()[-1])
print('$VAR=$v step ms', round(d['ms_per_step'],2), 'gemm', d['roofline']['kernels']['gemm_tc']['ms_per_call'])
"
done
