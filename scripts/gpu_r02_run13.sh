#!/bin/bash
# round 2, GPU call 13: racecheck / synccheck / initcheck of the AWACS kernel, racecheck of the general engine's models, and an
# ncu source-level capture of trial_kernel<MM1> (where do the general engine's instructions go?)
mkdir -p gpurun_out
for tool in racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/racecheck_awacs.py 6 6 > gpurun_out/r02_run13_awacs_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r02_run13_awacs_$tool.log; tail -3 gpurun_out/r02_run13_awacs_$tool.log
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:trial_kernel -c 1 -o gpurun_out/r02_engine_mm1 \
   python scripts/ncu_model.py 0 16384 4000 1 1.1111111 1.0 16 > gpurun_out/r02_run13_ncu.log 2>&1
tail -3 gpurun_out/r02_run13_ncu.log
ncu -i gpurun_out/r02_engine_mm1.ncu-rep --page source --csv --print-source cuda > gpurun_out/r02_engine_mm1_source.csv 2>/dev/null
ncu -i gpurun_out/r02_engine_mm1.ncu-rep --page raw --csv > gpurun_out/r02_engine_mm1_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
