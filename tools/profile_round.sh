#!/bin/bash
# tools/profile_round.sh TAG -- run on the GPU box (one GPU): ncu launch list + one `--set full` capture per BASELINE config of the
# fused kernel (and of the consumer kernel on c2), summarised into gpurun_out/TAG_*; compute-sanitizer logs of tools/sanitize_cases.py.
tag=${1:-r2}
B="python bench.py --no-e2e --no-cpu --no-decode --sustain-seconds 0 --no-check"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_c2.csv $B --no-consumer --steps 3 --warmup 3 > /dev/null 2>&1
for cfg in c2 c3_444 c3_gray c4 c5; do
  ncu --set full --clock-control none --import-source on -k regex:'fast_kernel' -s 2 -c 1 -f -o gpurun_out/${tag}_ncu_full_$cfg $B --no-consumer --config $cfg --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_$cfg.log 2>&1
  python tools/ncu_summary.py gpurun_out/${tag}_ncu_full_$cfg.ncu-rep gpurun_out/${tag}_ncu_full_$cfg > /dev/null 2>&1
  python tools/sass_mix.py gpurun_out/${tag}_ncu_full_$cfg.ncu-rep gpurun_out/${tag}_ncu_full_${cfg}_mix.txt > /dev/null 2>&1
  [ $cfg = c2 ] || rm -f gpurun_out/${tag}_ncu_full_$cfg.ncu-rep
done
ncu --set full --clock-control none --import-source on -k regex:'convert_kernel' -s 1 -c 1 -f -o gpurun_out/${tag}_ncu_full_consumer $B --steps 2 --warmup 1 > gpurun_out/${tag}_ncu_consumer.log 2>&1
python tools/ncu_summary.py gpurun_out/${tag}_ncu_full_consumer.ncu-rep gpurun_out/${tag}_ncu_full_consumer > /dev/null 2>&1
rm -f gpurun_out/${tag}_ncu_full_consumer.ncu-rep
# (racecheck: the fused-kernel cases only -- with the consumer / multi-device / strip-pipeline cases it ran into the 900 s limit)
for tool in ${SANITIZERS-memcheck racecheck synccheck}; do
  [ $tool = racecheck ] && export ZJ_SANITIZE_FUSED_ONLY=1 || unset ZJ_SANITIZE_FUSED_ONLY
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_cases.py > gpurun_out/${tag}_sanitizer_$tool.log 2>&1
  tail -3 gpurun_out/${tag}_sanitizer_$tool.log
done
ls -la gpurun_out/${tag}_* | head -40
