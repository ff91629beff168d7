"""AWACS (MODEL_AWACS) on the GPU against the plain-C oracle: same terrain, same seeds.

    python scripts/awacs_check.py [--trials 8] [--seconds 180] [--width 12 --height 10] [--out gpurun_out/awacs_check.json]

Prints, per trial, whether events / targets found / per-target detect states / the pop trace are identical, the
first diverging pop if not, and the aggregate detection statistics of both sides; then times a larger batch.
The oracle (tests/oracle_libs.py) is the checker only."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import cimba_b200 as cb                                             # noqa: E402
from oracle_libs import AWACS_TERRAIN_SEED, awacs_terrain, awacs_trial, load_port   # noqa: E402

MASTER = 0x34F05C64D7AD598F


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=8)
    ap.add_argument("--seconds", type=int, default=180)
    ap.add_argument("--width", type=float, default=12.0)
    ap.add_argument("--height", type=float, default=10.0)
    ap.add_argument("--batch", type=int, default=592)
    ap.add_argument("--trace", type=int, default=4000)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    assert (args.seconds / 3600.0) * 3600.0 == args.seconds
    port = load_port()
    ter = awacs_terrain(port, "port", AWACS_TERRAIN_SEED, args.width, args.height)
    m, cols, rows, geom = ter
    dev = torch.device("cuda", 0)
    cb.awacs_set_terrain(torch.from_numpy(m).to(dev), cols, rows, geom)
    res, per = cb.awacs_run(args.trials, duration_s=args.seconds, master_seed=MASTER, trace_cap=args.trace, device=dev)
    ev = res.events.cpu().numpy(); found = res.objects.cpu().numpy(); st = res.status.cpu().numpy()
    tds = per["tds"].cpu().numpy(); xs = per["x"].cpu().numpy()
    tk = res.trace_key.cpu().numpy(); tt = res.trace_time.cpu().numpy()
    report = {"trials": [], "seconds": args.seconds, "grid": [cols, rows]}
    exact = 0
    for i in range(args.trials):
        o, keys, times, p = awacs_trial(port, "port", cb.fmix64(MASTER, i), args.seconds / 3600.0, ter, trace_cap=args.trace)
        n = min(len(keys), args.trace, int(ev[i]))
        same_trace = list(tk[i][:n]) == keys[:n] and list(tt[i][:n]) == times[:n]
        first = None
        if not same_trace:
            for j in range(n):
                if tk[i][j] != keys[j] or tt[i][j] != times[j]:
                    first = [j, int(tk[i][j]), float(tt[i][j]), keys[j], times[j]]
                    break
        same = (int(ev[i]) == o.events and int(found[i]) == o.num_found and list(tds[i]) == p["tds"]
                and np.array_equal(xs[i].view(np.uint32), p["x"].view(np.uint32)) and same_trace)
        exact += bool(same)
        row = {"trial": i, "exact": bool(same), "events": [int(ev[i]), o.events], "found": [int(found[i]), o.num_found],
               "tds_differs": int(sum(a != b for a, b in zip(tds[i], p["tds"]))), "status": int(st[i]),
               "x_differs": int((xs[i].view(np.uint32) != p["x"].view(np.uint32)).sum()),
               "tds_gpu": np.bincount(tds[i], minlength=6).tolist(), "tds_oracle": list(o.tds_count), "first_diff": first}
        report["trials"].append(row)
        print(json.dumps(row), flush=True)
    report["exact"] = exact
    print(f"exact {exact} of {args.trials}", flush=True)
    # throughput of a batch that fills the machine (148 SMs x 4 warps per CTA)
    if args.batch:
        torch.cuda.synchronize()
        t0 = time.time()
        r2, _ = cb.awacs_run(args.batch, duration_s=args.seconds, master_seed=MASTER, device=dev)
        dt = time.time() - t0
        tot = int(r2.events.sum().item())
        report["batch"] = {"trials": args.batch, "wall_s": dt, "events": tot, "events_per_s": tot / dt,
                           "target_sweeps_per_s": args.batch * args.seconds * 1000 / dt,
                           "found_mean": float(r2.objects.double().mean().item())}
        print(json.dumps(report["batch"]), flush=True)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
