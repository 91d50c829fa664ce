#!/bin/bash
# tools/sweep.sh TAG "variant:spc variant:spc ..." [bench args] -- device-resident bench of tuning builds (run on the GPU box)
tag=$1; shift; list=$1; shift
for it in $list; do
  v=${it%%:*}; s=${it##*:}
  if [ "$s" = d ]; then unset ZJ_SPC; else export ZJ_SPC=$s; fi   # d = the launcher's own choice
  ZJ_LIB_PATH=build/variants/libzj_$v.so python bench.py --no-e2e --no-cpu --no-decode --sustain-seconds 0 --steps 10 "$@" > gpurun_out/${tag}_${v}_$s.json 2> gpurun_out/${tag}_${v}_$s.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_${v}_$s.json").read().strip().splitlines()[-1]); print("$v spc=$s", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["checked_vs_oracle"])
except Exception as e: print("$v spc=$s ERR", e, open("gpurun_out/${tag}_${v}_$s.err").read()[-400:])
PY
done
