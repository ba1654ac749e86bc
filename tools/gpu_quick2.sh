#!/bin/bash
# quick GPU check: selected tests + sampler microbench + short C2 bench
cd "$(dirname "$0")/.."
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -q -x -s -p no:cacheprovider -k "diagnostics or loop_control or dropin_sampler" > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python tools/microbench.py --only samp > gpurun_out/${TAG}_mb_samp.txt 2>&1; tail -5 gpurun_out/${TAG}_mb_samp.txt
timeout 600 python bench.py --steps 3 --no-cpu --no-gpu-ref --no-modes > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k:(v.get('launches_per_call'), v.get('ms_per_call')) for k,v in d['roofline']['kernels'].items()})
"
