"""Throughput of MODEL_AWACS (tutorial/tut_5_1.c) on one GPU at a chosen scale, CUDA-event timed.

    python scripts/awacs_bench.py --width 100 --height 100 --seconds 600 --trials 592 [--reps 2]

The terrain comes from the plain-C oracle's generator (test infrastructure used as a data source only: the map is
model INPUT, built by user code in the tutorial too); the simulation itself runs on the device.  Reports events/s,
target sweeps/s (targets x radar ticks) and terrain look-ups/s of the line-of-sight marches with their algorithmic
bytes (4 B each)."""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cimba_b200 as cb                                             # noqa: E402
from oracle_libs import AWACS_TERRAIN_SEED, awacs_terrain, load_port   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=float, default=100.0)
    ap.add_argument("--height", type=float, default=100.0)
    ap.add_argument("--seconds", type=int, default=600)
    ap.add_argument("--trials", type=int, default=592)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1, help="host threads for the terrain generator")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="also time this many trials of the unmodified reference model on all host cores (oracle/_ref)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    m, cols, rows, geom = awacs_terrain(load_port(), "port", AWACS_TERRAIN_SEED, args.width, args.height, args.threads)
    dev = torch.device("cuda", 0)
    cb.awacs_set_terrain(torch.from_numpy(m).to(dev), cols, rows, geom)
    cb.awacs_run(8, duration_s=30, master_seed=1, device=dev)              # warm-up
    best = None
    for rep in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res, _ = cb.awacs_run(args.trials, duration_s=args.seconds, master_seed=0x34F05C64D7AD598F,
                              first_trial=rep * args.trials, device=dev)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        events = int(res.events.sum().item())
        lookups = int(res.counters[:, 7].sum().item())
        row = {"grid": [cols, rows], "map_MB": cols * rows * 4 / 1e6, "trials": args.trials, "seconds": args.seconds,
               "ms": ms, "events_per_s": events / ms * 1e3, "target_sweeps_per_s": args.trials * args.seconds * 1000 / ms * 1e3,
               "lookups": lookups, "lookups_per_s": lookups / ms * 1e3, "lookup_GBps_algorithmic": lookups * 4 / ms / 1e6,
               "found_mean": float(res.objects.double().mean().item()), "bad": int((res.status != 0).sum().item())}
        print(json.dumps(row), flush=True)
        if best is None or row["ms"] < best["ms"]:
            best = row
    if args.cpu_sample:
        import ctypes as C
        import time
        from oracle_libs import awacs_ref_experiment, load_awacs_ref
        ref = load_awacs_ref()
        if ref is None:
            best["cpu_baseline"] = {"unavailable": "oracle/_ref/libawacs_ref.so is not built"}
        else:
            ref.awacs_ref_adopt_terrain(m.ctypes.data_as(C.POINTER(C.c_float)), C.c_uint32(cols), C.c_uint32(rows),
                                        geom.ctypes.data_as(C.POINTER(C.c_float)))
            t0 = time.time()
            outs = awacs_ref_experiment(ref, 0x34F05C64D7AD598F, 0, args.cpu_sample, args.seconds / 3600.0)
            dt = time.time() - t0
            best["cpu_baseline"] = {"kind": "reference", "cores": os.cpu_count(), "sample": f"{args.cpu_sample} trials x {args.seconds} s",
                                    "seconds": dt, "target_sweeps_per_s": args.cpu_sample * args.seconds * 1000 / dt,
                                    "events_per_s": sum(o.events for o in outs) / dt,
                                    "found_mean": sum(o.num_found for o in outs) / len(outs)}
        print(json.dumps(best["cpu_baseline"]), flush=True)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(best, indent=1))


if __name__ == "__main__":
    main()
