#!/bin/bash
# round 2, GPU call 16: full GPU suite with the static tier, ncu of static_trial_kernel, bench.py (default run, new secondaries)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/r02_run16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run16_pytest.log
tail -9 gpurun_out/r02_run16_pytest.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:static_trial_kernel -c 1 -o gpurun_out/r02_static_mm1 \
   python scripts/ncu_model.py 0 65536 4000 1 1.1111111 1.0 17 > gpurun_out/r02_run16_ncu.log 2>&1
tail -2 gpurun_out/r02_run16_ncu.log
ncu -i gpurun_out/r02_static_mm1.ncu-rep --page raw --csv > gpurun_out/r02_static_mm1_raw.csv 2>/dev/null
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_run16.json 2> gpurun_out/r02_bench_1gpu_run16.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_1gpu_run16.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
for s in d.get('secondary', []):
    print(' ', s.get('workload', s)[:90], '%.4g' % s.get('value', 0), s.get('failed_trials'), s.get('parity', {}).get('bit_identical'))
PY
