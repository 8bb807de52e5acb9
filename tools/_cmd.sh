mkdir -p gpurun_out
{
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "" "0,2"; do echo "== prep_only IFK_PREP_CFG=$cfg"; IFK_PREP_CFG=$cfg python tools/prep_only.py 2>&1 | tail -3; done
for cfg in "" "0,2"; do
echo "== step_breakdown IFK_PREP_CFG=$cfg"; IFK_PREP_CFG=$cfg timeout 200 python tools/step_breakdown.py 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k: round(v,4) for k,v in d.items() if k in ('forward_ms','forward_solves_only_ms','prepares_only_ms','backward_full_ms','step_ms','images_per_s')})"
done
} > gpurun_out/prep2.log 2>&1
cat gpurun_out/prep2.log
