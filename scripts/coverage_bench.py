"""The reference's own test worlds (models 3-6, 8, 11-14) on the general engine (CIMBA_B200_VARIANT_GENERAL) next to their
fixed-capacity kernels (variant 0: csrc/general.cuh + the engine as repair pass): same trials, same answers, CUDA-event timed.

    python scripts/coverage_bench.py [--trials 32768] [--duration 200] [--out gpurun_out/coverage_bench.json]"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import cimba_b200 as cb     # noqa: E402
from engine_bench import MASTER, timed     # noqa: E402

WORLDS = (("guarded objectqueue (test_objectqueue.c)", cb.MODEL_GUARDED, 10, 1.0, 1.0),
          ("... with its history", cb.MODEL_GUARDED_RECORDED, 10, 1.0, 1.0),
          ("priorityqueue (test_priorityqueue.c)", cb.MODEL_PRIOQ_RECORDED, 10, 1.0, 1.0),
          ("pool with pre-emption", cb.MODEL_PREEMPT, 20, 1.0, 1.0),
          ("buffer + resource", cb.MODEL_BUFFER, 10, 1.0, 1.0),
          ("buffer (test_buffer.c)", cb.MODEL_BUFFER_RECORDED, 10, 1.0, 1.0),
          ("priority queue by handle + condition", cb.MODEL_PRIOQ, 8, 1.0, 1.0),
          ("timers, waits, observers", cb.MODEL_TIMERS, 1, 1.0, 0.6),
          ("resource (test_resource.c)", cb.MODEL_RESOURCE_RECORDED, 1, 1.0, 1.0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=32768)
    ap.add_argument("--duration", type=int, default=200)
    ap.add_argument("--out", default="")
    ap.add_argument("--harbor-trials", type=int, default=4096)
    ap.add_argument("--harbor-hours", type=int, default=600)
    ap.add_argument("--harbor-only", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    for name, model, servers, arr, srv in (() if a.harbor_only else WORLDS):
        am = torch.full((a.trials,), arr, dtype=torch.float64, device=dev)
        sm = torch.full((a.trials,), srv, dtype=torch.float64, device=dev)
        out = {}
        for label, variant in (("engine", cb.VARIANT_GENERAL), ("round1_kernel", 0)):
            bufs = cb.TrialBuffers(a.trials, dev, 0, model, servers, variant)
            cb.launch_trials(am[:256], sm[:256], num_objects=20, master_seed=1, model=model, servers=servers, variant=variant)
            res, ms = timed(lambda: cb.launch_trials(am, sm, num_objects=a.duration, master_seed=MASTER, model=model, servers=servers,
                                                     variant=variant, buffers=bufs))
            ev = int(res.events.sum().item())
            out[label] = {"ms": ms, "events_per_s": ev / ms * 1e3, "events": ev, "bad": int((res.status != 0).sum().item()),
                          "sum_check": float(res.sum_wait.sum().item()), "counter_check": int(res.counters.sum().item()),
                          "workspace_MB": bufs.workspace_bytes / 1e6}
        row = {"world": name, "model": model, "trials": a.trials, "duration": a.duration, **out,
               "engine_over_round1_time": out["engine"]["ms"] / out["round1_kernel"]["ms"],
               "same_answers": all(out["engine"][k] == out["round1_kernel"][k] for k in ("events", "sum_check", "counter_check"))}
        rows.append(row)
        print(json.dumps(row), flush=True)
    # the harbor (test/test_condition.c): its fused kernels (warp per trial on chip up to a few thousand trials, lane per trial in
    # HBM beyond) next to harbor_general_model.cuh on the general engine
    for trials in (a.harbor_trials, 16 * a.harbor_trials):
        am = torch.full((trials,), 2.0, dtype=torch.float64, device=dev)
        sm = torch.full((trials,), 8.0, dtype=torch.float64, device=dev)
        out = {}
        for label, variant in (("engine", cb.VARIANT_GENERAL), ("fused", 0)):
            bufs = cb.TrialBuffers(trials, dev, 0, cb.MODEL_HARBOR, 10, variant)
            cb.launch_trials(am[:64], sm[:64], num_objects=50, master_seed=1, model=cb.MODEL_HARBOR, servers=10, variant=variant)
            res, ms = timed(lambda: cb.launch_trials(am, sm, num_objects=a.harbor_hours, master_seed=MASTER, model=cb.MODEL_HARBOR, servers=10,
                                                     variant=variant, buffers=bufs))
            ev = int(res.events.sum().item())
            out[label] = {"ms": ms, "events_per_s": ev / ms * 1e3, "events": ev, "bad": int((res.status != 0).sum().item()),
                          "sum_check": float(res.sum_wait.sum().item()), "counter_check": int(res.counters.sum().item())}
            del bufs
        row = {"world": "harbor (test_condition.c)", "model": cb.MODEL_HARBOR, "trials": trials, "duration": a.harbor_hours, **out,
               "engine_over_fused_time": out["engine"]["ms"] / out["fused"]["ms"],
               "same_answers": all(out["engine"][k] == out["fused"][k] for k in ("events", "sum_check", "counter_check"))}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
