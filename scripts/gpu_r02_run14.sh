#!/bin/bash
# round 2, GPU call 14: the static tier on the device - parity tests, then M/M/1 fused vs static vs general
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cmb_engine.py -q -x -k "static or user" > gpurun_out/r02_run14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run14_pytest.log
tail -15 gpurun_out/r02_run14_pytest.log
timeout 600 python scripts/engine_bench.py --out gpurun_out/r02_engine_bench_static.json > gpurun_out/r02_run14_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/r02_run14_bench.log
head -c 1500 gpurun_out/r02_run14_bench.log
