#!/bin/bash
# round 2, GPU call 2: the general engine on the device (tests + throughput), AWACS chunk / register sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cmb_engine.py -x -q > gpurun_out/r02_run2_pytest_engine.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run2_pytest_engine.log
tail -25 gpurun_out/r02_run2_pytest_engine.log
timeout 600 python scripts/engine_bench.py --out gpurun_out/r02_engine_bench.json > gpurun_out/r02_run2_engine_bench.log 2>&1
cat gpurun_out/r02_run2_engine_bench.log | tail -8
timeout 600 bash scripts/awacs_sweep.sh > gpurun_out/r02_run2_awacs_sweep.log 2>&1
cat gpurun_out/r02_run2_awacs_sweep.log
