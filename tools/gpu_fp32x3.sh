#!/bin/bash
# fp32x3 mode with the tensor-core split attention: parity tests, then the C2 step (and the CUDA-core kernel for comparison)
cd "$(dirname "$0")/.."
TAG=${1:-x3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_precision.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest precision rc=$?"; grep -E "tensor-core vs|passed|failed|Error" gpurun_out/${TAG}_tests.log | tail -8
for m in "" cuda; do
DV_ATTN_SPLIT=$m timeout 900 python bench.py --precision fp32x3 --steps 2 --no-cpu --no-gpu-ref --no-modes > gpurun_out/${TAG}_bench_fp32x3_$m.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_fp32x3_$m.json').read().strip().splitlines()[-1])
print('DV_ATTN_SPLIT=$m step ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:(v.get('launches_per_call'), v.get('ms_per_call')) for k,v in d['roofline']['kernels'].items()})
"
done
