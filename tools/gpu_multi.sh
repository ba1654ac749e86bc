#!/bin/bash
# multi-GPU bench lines (one box, N GPUs, torchrun/NCCL): usage gpu_multi.sh N tag "workloads"
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-mg}; WL=${3:-"c4"}
mkdir -p gpurun_out
for w in $WL; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --steps 3 --warmup 3 --no-cpu --no-gpu-ref --no-modes > gpurun_out/${TAG}_bench_${w}_n$N.json 2> gpurun_out/${TAG}_bench_${w}_n$N.err; echo "bench $w N=$N rc=$?"
tail -2 gpurun_out/${TAG}_bench_${w}_n$N.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_${w}_n$N.json').read().strip().splitlines()[-1])
print('$w N=$N', d['scaling'], 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['config'].get('B_per_gpu'), d['config'].get('global_batch'))
"
done
