#!/bin/bash
# Final-build measurement set of the round (one GPU): parity suite, bench lines, phase probes, step breakdown,
# ncu launch list of the bench command and one full capture of the dominant kernel.
mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
echo "== bench"; timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d["roofline"]["frac"], d["roofline"]["terms_us"], d["cpu_baseline"]["value"], d.get("parity"))
print({k: (v.get("value") if isinstance(v, dict) else v) for k, v in d.get("workloads", {}).items()})
PY
for dims in "100 12 16 16 3 1" "100 24 8 8 3 1" "100 48 4 4 3 1" "256 12 16 16 3 1" "256 24 8 8 3 1"; do
  timeout 120 python tools/probe_solve.py $dims 2>&1 | grep -v "^  File"
done > gpurun_out/solve_phase_probe.txt 2>&1
timeout 120 python tools/probe_solve.py 100 12 16 16 3 1 --chain 4 --no-parity 2>&1 | grep "unit" >> gpurun_out/solve_phase_probe.txt
cat gpurun_out/solve_phase_probe.txt | grep "per launch\|unit"
for wl in glow_imagenet32 glow_cifar glow_mnist; do timeout 200 python tools/step_breakdown.py --workload $wl 2>&1 | tail -1; done > gpurun_out/step_breakdown.txt
cat gpurun_out/step_breakdown.txt | cut -c1-400
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_in32.csv python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_wave -s 2 -c 1 -o gpurun_out/wave_100x12x16 -f python tools/profile_one.py 100 12 16 16 3 1 inverse 4 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -30
} > gpurun_out/final_measure.log 2>&1
tail -45 gpurun_out/final_measure.log
