#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"
grep -E "rel |passed|failed|FAILED|Error" gpurun_out/t_all.log | head -40
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench_c2.json; tail -n 5 gpurun_out/bench_c2.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/smoke.log
