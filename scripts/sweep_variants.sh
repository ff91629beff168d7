#!/bin/bash
# usage: scripts/sweep_variants.sh  (on the GPU box) - events/s of every library build under cimba_b200/lib/variants
for so in cimba_b200/lib/variants/*.so; do
  CIMBA_B200_LIB=$PWD/$so python bench.py --objects 100000 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$so', '%.4g' % d['value'], 'ms %.2f' % d['ms_per_step'], 'failed', d['failed_trials'], 'mean', d['summary']['mean_time_in_system'])"
done
