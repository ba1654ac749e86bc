#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_unet.py -q -m gpu -x > gpurun_out/t_unet_lanes.log 2>&1; echo "unet rc=$?"; tail -n 3 gpurun_out/t_unet_lanes.log
for l in 1 2 3 4; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --lanes $l > gpurun_out/bench_lanes$l.json 2> gpurun_out/bench_lanes$l.err; echo "lanes $l rc=$?"
tail -n 2 gpurun_out/bench_lanes$l.err
python -c "
import json; d=json.load(open('gpurun_out/bench_lanes$l.json')); print('lanes $l value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"
done
