#!/bin/bash
# round 2, GPU call 25: the whole GPU suite, memcheck over every kernel family, bench.py default run
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --durations=6 > gpurun_out/r02_run25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run25_pytest.log
tail -10 gpurun_out/r02_run25_pytest.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r02_run25_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_run25_memcheck.log
tail -3 gpurun_out/r02_run25_memcheck.log
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_run25.json 2> gpurun_out/r02_bench_1gpu_run25.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_1gpu_run25.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value'])
for s in d.get('secondary', []):
    print(' ', s.get('workload', str(s))[:80], '%.4g' % s.get('value', 0), s.get('failed_trials'), s.get('parity', {}).get('bit_identical'))
PY
