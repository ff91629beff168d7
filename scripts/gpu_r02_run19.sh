#!/bin/bash
# round 2, GPU call 19: static tier with packed keys + compile-time kinds: tests, timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cmb_engine.py -q -x -k "static or user" 2>&1 | tail -3
timeout 300 python scripts/engine_bench.py --static-only 2>&1 | tail -1
timeout 600 python scripts/engine_bench.py --gg1 2>&1 | tail -1
