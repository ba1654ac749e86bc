#!/bin/bash
# One gpurun call (1 GPU): the whole GPU test suite, smoke(), the bench lines of round 2 and the ncu evidence.
cd "$(dirname "$0")/.."
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; grep smoke: gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_c2_n1.json 2> gpurun_out/${TAG}_bench_c2.err; echo "bench c2 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_c2_reference_arm.json 2>/dev/null; echo "ref arm rc=$?"
for w in c1 c3 c5; do timeout 600 python bench.py --workload $w --no-modes > gpurun_out/${TAG}_bench_$w.json 2>/dev/null; echo "bench $w rc=$?"; done
timeout 600 python bench.py --workload c4 --steps 2 --no-cpu --no-gpu-ref --no-modes > gpurun_out/${TAG}_bench_c4_n1.json 2>/dev/null; echo "bench c4 (256 on one GPU) rc=$?"
# ncu: every launch of one eager denoiser call (C2 shapes): device time + DRAM bytes per launch (cold cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${TAG}_ncu_launches_one_call_c2.csv \
  --nvtx --nvtx-include "dv_call/" python tools/profile_call.py --workload c2 --calls 1 --no-events > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
# ncu --set full: first resnet + first transformer block at level 0, and a level-1 transformer block (tcgen05 attention)
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dv_call/" --launch-skip 3 --launch-count 22 \
  -o gpurun_out/${TAG}_prof_l0 -f python tools/profile_call.py --workload c2 --calls 1 --no-events > gpurun_out/${TAG}_ncu_l0.log 2>&1; echo "ncu l0 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "dv_call/" -k regex:attention_tc --launch-count 2 \
  -o gpurun_out/${TAG}_prof_attn_tc -f python tools/profile_call.py --workload c2 --calls 1 --no-events > gpurun_out/${TAG}_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:axpbypcz --launch-skip 25 --launch-count 2 \
  -o gpurun_out/${TAG}_prof_sampler -f python tools/microbench.py --only samp > gpurun_out/${TAG}_ncu_sampler.log 2>&1; echo "ncu sampler rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
