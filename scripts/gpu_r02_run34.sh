#!/bin/bash
# round 2, GPU call 34: final tree - the whole GPU suite, smoke(), bench.py default run
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r02_run34_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run34_pytest.log
tail -4 gpurun_out/r02_run34_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_run34.json 2> gpurun_out/r02_bench_1gpu_run34.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_1gpu_run34.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['gpu_launches'])
for s in d.get('secondary', []):
    print(' ', s.get('workload', str(s))[:70], '%.4g' % s.get('value', 0), s.get('failed_trials'), s.get('parity', {}).get('bit_identical'))
PY
