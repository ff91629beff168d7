#!/bin/bash
# round 2, GPU call 11: the whole GPU suite with the coverage models on the general engine, then engine vs round-1 kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > gpurun_out/r02_run11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run11_pytest.log
tail -15 gpurun_out/r02_run11_pytest.log
timeout 600 python scripts/coverage_bench.py --out gpurun_out/r02_coverage_bench.json > gpurun_out/r02_run11_coverage.log 2>&1; echo "bench rc=$?" >> gpurun_out/r02_run11_coverage.log
tail -12 gpurun_out/r02_run11_coverage.log | cut -c1-400
