#!/bin/bash
# tools/e2e_sweep.sh "streams:budgetMB ..." -- end-to-end (host buffers) bench for staging configurations (run on the GPU box)
for it in $1; do
  s=${it%%:*}; b=${it##*:}
  ZJ_E2E_STREAMS=$s ZJ_E2E_BUDGET_MB=$b python bench.py --no-cpu --no-decode --no-check --steps 3 --warmup 3 > gpurun_out/e2e_${s}_$b.json 2> gpurun_out/e2e_${s}_$b.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e2e_${s}_$b.json").read().strip().splitlines()[-1]); print("streams=$s budget=$b", d["e2e"]["value"], d["e2e"]["ms_per_step"])
except Exception as e: print("ERR", e)
PY
done
