#!/bin/bash
# One GPU call: parity suite, phase probes of the three model-shape solves, prepare variants, step breakdown, bench.
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for dims in "100 12 16 16 3 1" "100 24 8 8 3 1" "100 48 4 4 3 1" "256 12 16 16 3 1"; do
  echo "== probe $dims"; timeout 120 python tools/probe_solve.py $dims 2>&1 | grep -v "^  File"
done
for cfg in "" "1,0" "0,3" "0,9"; do
  echo "== step_breakdown IFK_PREP_CFG=$cfg"; IFK_PREP_CFG=$cfg timeout 200 python tools/step_breakdown.py 2>&1 | tail -1
done
echo "== bench"; timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench_default.json
} > gpurun_out/final_probe.log 2>&1
tail -60 gpurun_out/final_probe.log
