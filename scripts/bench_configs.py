#!/usr/bin/env python3
"""Secondary workloads of BASELINE.json (configs 3-4) and occupancy scaling of config 2.

Not the driver's bench (that is bench.py = config 2); this prints one JSON line per
configuration with device-resident throughput so DESIGN.md can quote measured numbers.
usage: python scripts/bench_configs.py [--objects N]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cimba_b200 as cb                                     # noqa: E402
from cimba_b200.experiment import TrialBuffers              # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--objects", type=int, default=100_000)
p.add_argument("--only", default="")
args = p.parse_args()

CONFIGS = [
    ("MM1 rho0.9 65536 trials (config 2)", cb.MODEL_MM1, 65536, 1 / 0.9, 1.0, 1),
    ("MM1 rho0.8 65536 trials (BASELINE.json rho)", cb.MODEL_MM1, 65536, 1.25, 1.0, 1),
    ("MM1 rho0.9 262144 trials (occupancy scaling)", cb.MODEL_MM1, 262144, 1 / 0.9, 1.0, 1),
    ("MM1 rho0.9 606208 trials (every warp slot busy)", cb.MODEL_MM1, 606208, 1 / 0.9, 1.0, 1),
    ("MMc c=8 rho0.8 32768 trials/GPU (config 3 shard)", cb.MODEL_MMC, 32768, 1 / 6.4, 1.0, 8),
    ("MMc c=8 rho0.8 262144 trials", cb.MODEL_MMC, 262144, 1 / 6.4, 1.0, 8),
    ("GG1 erlang2/normal rho0.8 65536 trials", cb.MODEL_GG1, 65536, 1.25, 1.0, 1),
    ("GG1 erlang2/normal rho0.8 1048576 trials (config 4)", cb.MODEL_GG1, 1048576, 1.25, 1.0, 1),
    ("HOLD 1000 processes 4096 trials (config 5's event-list shape; objects = duration/1000)", cb.MODEL_HOLD, 4096, 1.0, 1.0, 1000),
    ("HARBOR test_condition.c 4096 trials (config 5's cmb_condition shape; objects = hours/10)", cb.MODEL_HARBOR, 4096, 2.0, 8.0, 10),
    ("HARBOR test_condition.c 65536 trials", cb.MODEL_HARBOR, 65536, 2.0, 8.0, 10),
    ("GUARDED bounded queue under interrupts, 65536 trials (objects = duration/100)", cb.MODEL_GUARDED, 65536, 1.0, 1.0, 10),
    ("PREEMPT resourcepool with pre-emption, 65536 trials", cb.MODEL_PREEMPT, 65536, 1.0, 1.0, 20),
    ("BUFFER + resource, 65536 trials", cb.MODEL_BUFFER, 65536, 1.0, 1.0, 10),
    ("PRIOQ + condition, 65536 trials", cb.MODEL_PRIOQ, 65536, 1.0, 1.0, 8),
    ("TIMERS waits observers, 65536 trials", cb.MODEL_TIMERS, 65536, 1.0, 0.6, 1),
    ("HARBOR warp per trial, state in shared memory (variant 1), 4096 trials", cb.MODEL_HARBOR, 4096, 2.0, 8.0, 10, 1),
    ("HARBOR warp per trial, state in shared memory (variant 1), 65536 trials", cb.MODEL_HARBOR, 65536, 2.0, 8.0, 10, 1),
]
GENERAL = (cb.MODEL_GUARDED, cb.MODEL_PREEMPT, cb.MODEL_BUFFER, cb.MODEL_PRIOQ, cb.MODEL_TIMERS)
dev = torch.device("cuda", 0)
for name, model, n, arr, srv, servers, *rest in CONFIGS:
    variant = rest[0] if rest else 0
    if args.only and args.only not in name:
        continue
    a = torch.full((n,), arr, dtype=torch.float64, device=dev)
    s = torch.full((n,), srv, dtype=torch.float64, device=dev)
    bufs = TrialBuffers(n, dev, 0, model, servers, variant)
    size = args.objects // 1000 if model == cb.MODEL_HOLD else (args.objects // 10 if model == cb.MODEL_HARBOR else
                                                          (args.objects // 100 if model in GENERAL else args.objects))
    run = lambda: cb.launch_trials(a, s, num_objects=size, master_seed=0x34F05C64D7AD598F,
                                   model=model, servers=servers, buffers=bufs, variant=variant)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ev = res.total_events()
    print(json.dumps({"config": name, "trials": n, "objects": size, "events": ev, "ms": ms,
                      "events_per_s": ev / ms * 1e3, "failed": int((res.status != 0).sum()),
                      "mean_time_in_system": float((res.sum_wait / res.objects.double().clamp(min=1)).mean())}), flush=True)
    del bufs, a, s
