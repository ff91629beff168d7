#!/bin/bash
# usage: scripts/awacs_sweep.sh [awacs_bench.py arguments]   (on the GPU box)
# MODEL_AWACS throughput of the shipped library and of every build under cimba_b200/lib/variants
# (python scripts/build_variant.py awacs_chunk4 -DAWACS_CHUNK=4, ... built beforehand on the CPU box).
args=${@:---width 100 --height 100 --seconds 300 --trials 4096 --reps 1}
echo "shipped"; python scripts/awacs_bench.py $args 2>&1 | tail -1
for so in cimba_b200/lib/variants/awacs_*.so; do
  [ -e "$so" ] || continue
  echo "$so"; CIMBA_B200_LIB=$PWD/$so python scripts/awacs_bench.py $args 2>&1 | tail -1
done
