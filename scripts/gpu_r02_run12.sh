#!/bin/bash
# round 2, GPU call 12: the whole GPU suite, engine vs fixed-capacity kernels for the coverage models, memcheck over every kernel family
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r02_run12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run12_pytest.log
tail -15 gpurun_out/r02_run12_pytest.log
timeout 600 python scripts/coverage_bench.py --out gpurun_out/r02_coverage_bench.json > gpurun_out/r02_run12_coverage.log 2>&1; echo "bench rc=$?" >> gpurun_out/r02_run12_coverage.log
tail -3 gpurun_out/r02_run12_coverage.log | cut -c1-300
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > gpurun_out/r02_run12_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_run12_memcheck.log
tail -8 gpurun_out/r02_run12_memcheck.log
