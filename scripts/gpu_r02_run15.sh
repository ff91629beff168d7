#!/bin/bash
# round 2, GPU call 15: tuning sweep of the static tier (M/M/1, 65 536 trials x 1e5 objects)
mkdir -p gpurun_out
for so in cimba_b200/lib/variants/static_*.so; do
  echo -n "$so " >> gpurun_out/r02_static_sweep.log
  CIMBA_B200_LIB=$PWD/$so timeout 300 python scripts/engine_bench.py --static-only 2>&1 | tail -1 >> gpurun_out/r02_static_sweep.log
done
cat gpurun_out/r02_static_sweep.log
