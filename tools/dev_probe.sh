#!/bin/bash
run() { echo "== IFK_WAVE_CFG=$IFK_WAVE_CFG IFK_PDL=$IFK_PDL"; timeout 120 python tools/probe_solve.py "$@" 2>&1 | grep -v "bookkeeping\|constants\|tail  "; }
{
for cfg in "6,4,2" "6,8,2"; do export IFK_WAVE_CFG=$cfg; run 100 12 16 16 3 1; run 256 12 16 16 3 1 --no-parity; done
export IFK_WAVE_CFG=6,8,2; run 100 24 8 8 3 1;  run 256 24 8 8 3 1 --no-parity
export IFK_WAVE_CFG=6,16,2; run 100 48 4 4 3 1
unset IFK_WAVE_CFG
run 100 24 8 8 3 4; run 100 48 4 4 3 4; run 5 12 5 7 3 1
} | tee -a gpurun_out/probe.log
