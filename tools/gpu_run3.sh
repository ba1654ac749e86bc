#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
grep -E "rel |passed|failed|FAILED|Error" gpurun_out/t_all.log | head -70
timeout 300 python tools/profile_call.py --workload c2 --out gpurun_out/ops_c2.csv > gpurun_out/profile_c2.log 2>&1; echo "profile rc=$?"; head -n 3 gpurun_out/profile_c2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "dv_call/" --csv --log-file gpurun_out/launches_c2.csv python tools/profile_call.py --workload c2 --calls 1 --no-events > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu.log; wc -l gpurun_out/launches_c2.csv
