#!/bin/bash
# round 2, GPU call 37: after the second sfc64 change - recalibrate the issue roofline, the whole GPU suite, bench
mkdir -p gpurun_out
ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum \
    --clock-control none -k regex:'mm1_kernel|gg1_kernel|pool_fast_kernel' --csv --log-file gpurun_out/issue_ncu.csv \
    python scripts/calibrate_issue.py --run gpurun_out/issue_diag.json > gpurun_out/r02_run37_calib.log 2>&1; echo "calib rc=$?"
python scripts/calibrate_issue.py --combine gpurun_out/issue_diag.json gpurun_out/issue_ncu.csv > gpurun_out/r02_run37_combine.log 2>&1; tail -3 gpurun_out/r02_run37_combine.log
cp profiles/issue_calibration.json gpurun_out/issue_calibration.json
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r02_run37_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run37_pytest.log
tail -3 gpurun_out/r02_run37_pytest.log
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_run37.json 2> gpurun_out/r02_bench_1gpu_run37.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_1gpu_run37.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['warp_instructions_per_iteration'], d['clocks'])
for s in d.get('secondary', []):
    print(' ', s.get('workload', str(s))[:70], '%.4g' % s.get('value', 0), s.get('failed_trials'), s.get('parity', {}).get('bit_identical'))
PY
