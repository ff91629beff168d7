#!/bin/bash
# round 2, GPU call 17: sampled holds - engine tests, then G/G/1 fused vs static
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cmb_engine.py tests/test_gpu_parity.py -q -x -k "static or user or gg1 or GG1 or general_engine or erlang or normal" > gpurun_out/r02_run17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run17_pytest.log
tail -5 gpurun_out/r02_run17_pytest.log
timeout 600 python scripts/engine_bench.py --gg1 > gpurun_out/r02_run17_gg1.log 2>&1; tail -2 gpurun_out/r02_run17_gg1.log
timeout 300 python scripts/engine_bench.py --static-only 2>&1 | tail -1
