#!/bin/bash
# round 2, GPU call 38 (final tree): the calibration's raw ncu numbers, the launch list of a bench run, one --set full capture of mm1_kernel
mkdir -p gpurun_out
ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum \
    --clock-control none -k regex:'mm1_kernel|gg1_kernel|pool_fast_kernel' --csv --log-file gpurun_out/r02_issue_ncu_final.csv \
    python scripts/calibrate_issue.py --run gpurun_out/r02_issue_diag_final.json > /dev/null 2>&1; echo "calib rc=$?"
python scripts/calibrate_issue.py --combine gpurun_out/r02_issue_diag_final.json gpurun_out/r02_issue_ncu_final.csv > /dev/null 2>&1; cp profiles/issue_calibration.json gpurun_out/issue_calibration_final.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1; echo "launch list rc=$?"
ncu --set full --import-source on --clock-control none -k regex:mm1_kernel -c 1 -o gpurun_out/r02_mm1_kernel_final \
    python scripts/ncu_model.py 0 65536 20000 1 1.1111111 1.0 0 > /dev/null 2>&1; echo "full rc=$?"
ncu -i gpurun_out/r02_mm1_kernel_final.ncu-rep --page raw --csv > gpurun_out/r02_mm1_kernel_final_full_raw.csv 2>/dev/null
rm -f gpurun_out/r02_mm1_kernel_final.ncu-rep
ls -la gpurun_out/ | grep -E "final|launches"
