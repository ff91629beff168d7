#!/usr/bin/env python3
"""BASELINE.json configs 3 / 4 across the GPUs of one box: one process per GPU (torchrun), trials sharded
by global index, no data-path communication, per-GPU cmb_datasummary merged over NCCL at the end.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 \
      scripts/bench_multi.py --model mmc --trials 262144 --objects 1000000
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cimba_b200 as cb                                     # noqa: E402
from cimba_b200.experiment import TrialBuffers              # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--model", choices=["mm1", "mmc", "gg1"], default="mmc")
p.add_argument("--trials", type=int, default=262144, help="total over all GPUs")
p.add_argument("--objects", type=int, default=1_000_000)
p.add_argument("--steps", type=int, default=2)
args = p.parse_args()
MODEL = {"mm1": (cb.MODEL_MM1, 1 / 0.9, 1.0, 1), "mmc": (cb.MODEL_MMC, 1 / 6.4, 1.0, 8), "gg1": (cb.MODEL_GG1, 1.25, 1.0, 1)}[args.model]
model, arr, srv, servers = MODEL

sys.stdout.flush()
real_stdout = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)                                               # library banners go to stderr
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = args.trials // world
first = rank * n
a = torch.full((n,), arr, dtype=torch.float64, device=dev)
s = torch.full((n,), srv, dtype=torch.float64, device=dev)
bufs = TrialBuffers(n, dev, 0, model, servers)
step = lambda: cb.launch_trials(a, s, num_objects=args.objects, master_seed=0x34F05C64D7AD598F, first_trial=first,
                                model=model, servers=servers, buffers=bufs)
cb.launch_trials(a, s, num_objects=min(args.objects, 2000), master_seed=1, first_trial=first, model=model, servers=servers, buffers=bufs)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    res = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
t = torch.tensor([ms, float(res.total_events()), float((res.status != 0).sum())], dtype=torch.float64, device=dev)
per_rank = None
if world > 1:
    tmax, tsum = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    gathered = [None] * world
    dist.all_gather_object(gathered, float(t[0]))
    per_rank = gathered
    ms, events, bad = float(tmax[0]), float(tsum[1]), float(tsum[2])
else:
    events, bad = float(t[1]), float(t[2])
merged = cb.merge_across_ranks(cb.summarize_on_device(res.sum_wait, res.objects))
if rank == 0:
    real_stdout.write(json.dumps({"model": args.model, "n_gpus": world, "trials": n * world, "objects": args.objects,
                                  "events_per_step": events, "ms_per_step": ms, "events_per_s": events / (ms * 1e-3),
                                  "failed_trials": bad, "rank_ms_per_step": per_rank,
                                  "summary": {"n": merged.count(), "mean_time_in_system": merged.mean(),
                                              "ci95_half_width": merged.half_width_95()}}) + "\n")
    real_stdout.flush()
if world > 1:
    dist.destroy_process_group()
