#!/bin/bash
# round 2, GPU call 1: full GPU test suite, AWACS A/B sweep, ncu capture of the shipped awacs_kernel, headline bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_run1_gpu.txt
lscpu | head -20 > gpurun_out/r02_run1_lscpu.txt; nproc >> gpurun_out/r02_run1_lscpu.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02_run1_lscpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
tail -5 gpurun_out/r02_run1_pytest.log
timeout 600 bash scripts/awacs_sweep.sh > gpurun_out/r02_run1_awacs_sweep.log 2>&1
cat gpurun_out/r02_run1_awacs_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:awacs_kernel -s 1 -c 1 -o gpurun_out/r02_awacs_kernel_full \
    python scripts/awacs_bench.py --width 100 --height 100 --seconds 30 --trials 592 --reps 1 > gpurun_out/r02_run1_ncu_awacs.log 2>&1
tail -3 gpurun_out/r02_run1_ncu_awacs.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_run1_bench.json 2> gpurun_out/r02_run1_bench.err
cat gpurun_out/r02_run1_bench.json
