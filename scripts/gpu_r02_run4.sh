#!/bin/bash
# round 2, GPU call 4: the whole GPU suite on the shipped build, the general engine's second form, AWACS on the tutorial's full map
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r02_run4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run4_pytest.log
tail -14 gpurun_out/r02_run4_pytest.log
timeout 600 python scripts/engine_bench.py --out gpurun_out/r02_engine_bench.json > gpurun_out/r02_run4_engine_bench.log 2>&1
tail -5 gpurun_out/r02_run4_engine_bench.log | cut -c1-700
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trial_kernel -c 1 -o gpurun_out/r02_trial_kernel_mm1_v2_full \
    python -c "
import torch, cimba_b200 as cb
cb.run_trials(65536, arr_mean=1/0.9, srv_mean=1.0, num_objects=2000, master_seed=7, variant=3)
" > gpurun_out/r02_run4_ncu_engine.log 2>&1
timeout 1200 python scripts/awacs_full.py --width 1000 --height 1000 --hours 1 --trials 4096 --out gpurun_out/r02_awacs_fullmap_1h.json > gpurun_out/r02_run4_awacs_full_1h.log 2>&1; echo "awacs full 1h rc=$?"
tail -1 gpurun_out/r02_run4_awacs_full_1h.log | cut -c1-900
