#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-gs}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_precision.py -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
for prec in fp16s fp32x3; do
for f in 1 0; do
DV_GN_OUT=$f timeout 900 python bench.py --precision $prec --steps 2 --no-cpu --no-gpu-ref --no-modes > gpurun_out/${TAG}_bench_${prec}_gnout$f.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${prec}_gnout$f.json').read().strip().splitlines()[-1])
print('$prec DV_GN_OUT=$f step ms', round(d['ms_per_step'],2), {k:(v.get('launches_per_call'), v.get('ms_per_call')) for k,v in d['roofline']['kernels'].items() if 'launches_per_call' in v})
"
done
done
