#!/bin/bash
# fused output GroupNorm A/B over the workloads (DV_GN_OUT=1 default vs 0)
cd "$(dirname "$0")/.."
TAG=${1:-gno2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -m gpu -q -x -p no:cacheprovider -k "groupnorm_in_producer" > gpurun_out/${TAG}_tests1.log 2>&1; echo "pytest gn rc=$?"; tail -2 gpurun_out/${TAG}_tests1.log
for w in c2 c1 c3 c5; do
for f in 1 0; do
DV_GN_OUT=$f timeout 600 python bench.py --workload $w --steps 3 --no-cpu --no-gpu-ref --no-modes > gpurun_out/${TAG}_bench_${w}_gnout$f.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${w}_gnout$f.json').read().strip().splitlines()[-1])
print('$w DV_GN_OUT=$f step ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:(v.get('launches_per_call'), v.get('ms_per_call')) for k,v in d['roofline']['kernels'].items() if k in ('gemm_tc','gn_apply')})
"
done
done
