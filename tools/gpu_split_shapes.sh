#!/bin/bash
# split-precision GEMM tile shapes: parity tests, then fp16s / fp32x3 C2 step with the wide shapes (default) vs 128 x 128 only
cd "$(dirname "$0")/.."
TAG=${1:-ss}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_precision.py -m gpu -q -s -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest precision rc=$?"; grep -E "tensor-core vs|passed|failed|Error|fp16s t=|fp32x3 t=|dpm50|unipc10" gpurun_out/${TAG}_tests.log | tail -14
for prec in fp16s fp32x3; do
for m in "" 1; do
if [ -n "$m" ]; then export DV_GEMM_SPLIT128=1; else unset DV_GEMM_SPLIT128; fi
timeout 900 python bench.py --precision $prec --steps 2 --no-cpu --no-gpu-ref --no-modes > gpurun_out/${TAG}_bench_${prec}_$m.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${prec}_$m.json').read().strip().splitlines()[-1])
print('$prec SPLIT128=$m step ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:(v.get('launches_per_call'), v.get('ms_per_call')) for k,v in d['roofline']['kernels'].items() if 'launches_per_call' in v})
"
done
done
