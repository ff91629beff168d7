#!/bin/bash
# round 2, 8-GPU call: the bench line as the driver launches it at N = 8, then the single-process executive over the same 8 GPUs
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r02_8gpu_devices.txt; nproc >> gpurun_out/r02_8gpu_devices.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r02_8gpu_devices.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; echo "bench8 rc=$?"
head -c 1200 gpurun_out/r02_bench_8gpu.json; echo
timeout 600 python bench.py --gpus 8 --single-process --steps 5 --warmup 2 > gpurun_out/r02_bench_8gpu_single_process.json 2> gpurun_out/r02_bench_8gpu_single_process.err; echo "single rc=$?"
head -c 1500 gpurun_out/r02_bench_8gpu_single_process.json; echo
