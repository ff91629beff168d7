#!/bin/bash
# round 2, GPU call 3: everything once more on the shipped build - full GPU suite (with the 24-hour AWACS vectors), the new
# bench line (both arms), the issue calibration, AWACS 4096 x 24 h against the golden vectors, ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02_run3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run3_pytest.log
tail -14 gpurun_out/r02_run3_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_run3_bench.json 2> gpurun_out/r02_run3_bench.err; echo "bench rc=$?"
head -c 6000 gpurun_out/r02_run3_bench.json; tail -5 gpurun_out/r02_run3_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_run3_bench_ref.json 2> gpurun_out/r02_run3_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/r02_run3_bench_ref.json
timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum \
    --clock-control none -k regex:'mm1_kernel|gg1_kernel|pool_fast_kernel' --csv --log-file gpurun_out/issue_ncu.csv \
    python scripts/calibrate_issue.py --run gpurun_out/issue_diag.json > gpurun_out/r02_run3_calib.log 2>&1; echo "calib rc=$?"
cat gpurun_out/issue_diag.json
timeout 900 python scripts/awacs_full.py --width 100 --height 100 --hours 24 --trials 4096 --golden tests/golden/awacs_24h.npz --out gpurun_out/r02_awacs_24h.json > gpurun_out/r02_run3_awacs24.log 2>&1; echo "awacs24 rc=$?"
tail -2 gpurun_out/r02_run3_awacs24.log | cut -c1-1500
timeout 600 ncu --set full --clock-control none --import-source on -k regex:awacs_kernel -s 1 -c 1 -o gpurun_out/r02_awacs_kernel_4096_full \
    python scripts/awacs_bench.py --width 100 --height 100 --seconds 20 --trials 4096 --reps 1 > gpurun_out/r02_run3_ncu_awacs.log 2>&1
timeout 600 python scripts/engine_bench.py --out gpurun_out/r02_engine_bench.json > gpurun_out/r02_run3_engine_bench.log 2>&1
tail -5 gpurun_out/r02_run3_engine_bench.log | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trial_kernel -c 1 -o gpurun_out/r02_trial_kernel_mm1_full \
    python -c "
import torch, cimba_b200 as cb
cb.run_trials(16384, arr_mean=1/0.9, srv_mean=1.0, num_objects=5000, master_seed=7, variant=3)
" > gpurun_out/r02_run3_ncu_engine.log 2>&1
tail -2 gpurun_out/r02_run3_ncu_engine.log
