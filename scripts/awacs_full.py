"""BASELINE config 5 at (or towards) its definition: 4096 replications of tutorial/tut_5_1.c, 24-hour trials.

    python scripts/awacs_full.py --width 100 --height 100 --hours 24 --trials 4096 --golden tests/golden/awacs_24h.npz --out gpurun_out/r02_awacs_24h.json
    python scripts/awacs_full.py --width 1000 --height 1000 --hours 24 --trials 4096 --out gpurun_out/r02_awacs_24h_fullmap.json

One launch, CUDA-event timed.  With --golden (vectors the unmodified tutorial source produced, tests/golden/make_awacs_24h.py:
same map size, same seeds) the first trials are compared field by field: event count, end time, targets found, the detect-state
and mode counts and all 1000 final positions (float bits), modes, detect states and found flags."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cimba_b200 as cb                                             # noqa: E402
from oracle_libs import AWACS_TERRAIN_SEED, awacs_terrain, load_port   # noqa: E402

MASTER = 0x34F05C64D7AD598F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=float, default=100.0)
    ap.add_argument("--height", type=float, default=100.0)
    ap.add_argument("--hours", type=float, default=24.0)
    ap.add_argument("--trials", type=int, default=4096)
    ap.add_argument("--golden", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    t0 = time.time()
    threads = len(os.sched_getaffinity(0))
    m, cols, rows, geom = awacs_terrain(load_port(), "port", AWACS_TERRAIN_SEED, a.width, a.height, threads)
    t_gen = time.time() - t0
    dev = torch.device("cuda", 0)
    t0 = time.time()
    cb.awacs_set_terrain(torch.from_numpy(m).to(dev), cols, rows, geom)
    torch.cuda.synchronize()
    t_up = time.time() - t0
    del m
    seconds = int(round(a.hours * 3600.0))
    cb.awacs_run(8, duration_s=30, master_seed=1, device=dev)              # warm-up
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res, per = cb.awacs_run(a.trials, duration_s=seconds, master_seed=MASTER, device=dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    events = int(res.events.sum().item())
    cells = int(res.counters[:, 7].sum().item())
    row = {"grid": [cols, rows], "map_GB": cols * rows * 4 / 1e9, "terrain_generation_s": t_gen, "terrain_upload_s": t_up,
           "trials": a.trials, "simulated_hours": a.hours, "ms": ms, "events": events, "events_per_s": events / ms * 1e3,
           "target_sweeps_per_s": a.trials * seconds * 1000 / ms * 1e3, "cells_read": cells,
           "cells_read_per_s": cells / ms * 1e3, "cell_GBps_algorithmic": cells * 4 / ms / 1e6,
           "found_mean": float(res.objects.double().mean().item()), "events_mean": events / a.trials,
           "bad": int((res.status != 0).sum().item())}
    if a.golden:
        g = np.load(a.golden)
        n = min(len(g["events"]), a.trials)
        assert [cols, rows] == g["grid"].tolist() and float(g["hours"]) == a.hours
        cnt = res.counters.cpu().numpy()
        modes = np.stack([(cnt[:, 6] >> (16 * k)) & 0xffff for k in range(4)], axis=1)
        checks = {
            "events": res.events.cpu().numpy()[:n].tolist() == g["events"][:n].tolist(),
            "t_end": res.t_end.cpu().numpy()[:n].tolist() == g["t_end"][:n].tolist(),
            "num_found": res.objects.cpu().numpy()[:n].tolist() == g["num_found"][:n].tolist(),
            "sum_x": res.sum_wait.cpu().numpy()[:n].tolist() == g["sum_x"][:n].tolist(),
            "tds_count": cnt[:n, :6].tolist() == g["tds_count"][:n].tolist(),
            "mode_count": modes[:n].tolist() == g["mode_count"][:n].tolist(),
            "x_bits": bool(np.array_equal(per["x"][:n].cpu().numpy().view(np.uint32), g["x_bits"][:n])),
            "y_bits": bool(np.array_equal(per["y"][:n].cpu().numpy().view(np.uint32), g["y_bits"][:n])),
            "mode": bool(np.array_equal(per["mode"][:n].cpu().numpy().astype(np.uint8), g["mode"][:n])),
            "tds": bool(np.array_equal(per["tds"][:n].cpu().numpy().astype(np.uint8), g["tds"][:n])),
            "found": bool(np.array_equal(per["found"][:n].cpu().numpy().astype(np.uint8), g["det"][:n])),
        }
        row["golden"] = {"file": a.golden, "trials_compared": n, "checks": checks, "all_identical": all(checks.values()),
                         "events_gpu": res.events.cpu().numpy()[:n].tolist(), "events_reference": g["events"][:n].tolist()}
    print(json.dumps(row), flush=True)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(json.dumps(row, indent=1))


if __name__ == "__main__":
    main()
