#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-f16}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_precision.py -m gpu -q -s -p no:cacheprovider -k "fp16_mode" > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; grep -E "fp16 |passed|failed|Error" gpurun_out/${TAG}_tests.log | tail -14
for prec in fp16 bf16; do
timeout 900 python bench.py --precision $prec --steps 3 --no-cpu --no-gpu-ref --no-modes > gpurun_out/${TAG}_bench_${prec}.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${prec}.json').read().strip().splitlines()[-1])
print('$prec step ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k:(v.get('launches_per_call'), v.get('ms_per_call')) for k,v in d['roofline']['kernels'].items() if 'launches_per_call' in v})
"
done
