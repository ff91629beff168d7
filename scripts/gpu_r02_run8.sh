#!/bin/bash
# round 2, GPU call 8: general engine, fourth form (heap heads and the first process records inline) - tests and throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cmb_engine.py -x -q > gpurun_out/r02_run8_pytest_engine.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run8_pytest_engine.log
tail -4 gpurun_out/r02_run8_pytest_engine.log
timeout 600 python scripts/engine_bench.py --out gpurun_out/r02_engine_bench_v4.json > gpurun_out/r02_run8_engine_bench.log 2>&1
cat gpurun_out/r02_run8_engine_bench.log | cut -c1-420
