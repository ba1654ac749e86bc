#!/bin/bash
# fused output GroupNorm: parity tests, then A/B of the C2 step (DV_GN_OUT=1 default vs 0)
cd "$(dirname "$0")/.."
TAG=${1:-gno}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -q -x -s -p no:cacheprovider -k "groupnorm_in_producer" > gpurun_out/${TAG}_tests1.log 2>&1; echo "pytest gn rc=$?"; tail -4 gpurun_out/${TAG}_tests1.log
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest all rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
for f in 1 0; do
DV_GN_OUT=$f timeout 600 python bench.py --steps 3 --no-cpu --no-gpu-ref --no-modes > gpurun_out/${TAG}_bench_gnout$f.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_gnout$f.json').read().strip().splitlines()[-1])
print('DV_GN_OUT=$f step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k:(v.get('launches_per_call'), v.get('ms_per_call')) for k,v in d['roofline']['kernels'].items()})
"
done
