"""Throughput of the general engine (cimba_b200/csrc/cmb_device.cuh) next to the hand-fused kernels, CUDA-event timed.

    python scripts/engine_bench.py [--out gpurun_out/engine_bench.json]

M/M/1 and M/M/c: the model written against the authoring surface (CIMBA_B200_VARIANT_GENERAL) vs the fast kernel (variant 0), same
trials, same answers; the reneging model (1000 processes per trial) on its own."""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import cimba_b200 as cb     # noqa: E402

MASTER = 0x34F05C64D7AD598F


def timed(fn, reps=2):
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None or ms < best else best
    return res, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--pc-only", action="store_true", help="M/M/1: mm1_kernel next to mm1_pc_kernel (variates from producer warps)")
    ap.add_argument("--gg1", action="store_true", help="G/G/1: gg1_kernel next to gg1_model.cuh on the static tier")
    ap.add_argument("--static-only", action="store_true", help="time M/M/1 on the static tier only (variant sweeps)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    rows = []
    if a.gg1:
        trials, nobj = 262144, 100000
        am = torch.full((trials,), 1.25, dtype=torch.float64, device=dev)
        sm = torch.full((trials,), 1.0, dtype=torch.float64, device=dev)
        out = {}
        for label, variant in (("fast", 0), ("static", cb.VARIANT_STATIC)):
            bufs = cb.TrialBuffers(trials, dev, 0, cb.MODEL_GG1, 1, variant)
            cb.launch_trials(am[:256], sm[:256], num_objects=1000, master_seed=1, model=cb.MODEL_GG1, variant=variant)
            res, ms = timed(lambda: cb.launch_trials(am, sm, num_objects=nobj, master_seed=MASTER, model=cb.MODEL_GG1, variant=variant, buffers=bufs))
            ev = int(res.events.sum().item())
            out[label] = {"ms": ms, "events_per_s": ev / ms * 1e3, "events": ev, "bad": int((res.status != 0).sum().item()),
                          "sum_check": float(res.sum_wait.sum().item())}
            del bufs
        print(json.dumps({"model": "G/G/1", "trials": trials, "objects": nobj, **out,
                          "static_over_fast_time": out["static"]["ms"] / out["fast"]["ms"],
                          "same_answers": out["fast"]["sum_check"] == out["static"]["sum_check"] and out["fast"]["events"] == out["static"]["events"]}), flush=True)
        return
    if a.pc_only:
        trials, nobj = 65536, 100000
        am = torch.full((trials,), 1 / 0.9, dtype=torch.float64, device=dev)
        sm = torch.full((trials,), 1.0, dtype=torch.float64, device=dev)
        for label, variant in (("fast", 0), ("producer_consumer", 2)):
            bufs = cb.TrialBuffers(trials, dev, 0, cb.MODEL_MM1, 1, variant)
            cb.launch_trials(am[:256], sm[:256], num_objects=1000, master_seed=1, variant=variant)
            res, ms = timed(lambda: cb.launch_trials(am, sm, num_objects=nobj, master_seed=MASTER, variant=variant, buffers=bufs), reps=3)
            ev = int(res.events.sum().item())
            print(json.dumps({"kernel": label, "ms": ms, "events_per_s": ev / ms * 1e3, "bad": int((res.status != 0).sum().item()),
                              "sum_check": float(res.sum_wait.sum().item())}), flush=True)
        return
    if a.static_only:
        trials, nobj = 65536, 100000
        am = torch.full((trials,), 1 / 0.9, dtype=torch.float64, device=dev)
        sm = torch.full((trials,), 1.0, dtype=torch.float64, device=dev)
        bufs = cb.TrialBuffers(trials, dev, 0, cb.MODEL_MM1, 1, cb.VARIANT_STATIC)
        cb.launch_trials(am[:256], sm[:256], num_objects=1000, master_seed=1, variant=cb.VARIANT_STATIC)
        res, ms = timed(lambda: cb.launch_trials(am, sm, num_objects=nobj, master_seed=MASTER, variant=cb.VARIANT_STATIC, buffers=bufs), reps=3)
        ev = int(res.events.sum().item())
        print(json.dumps({"static_ms": ms, "events_per_s": ev / ms * 1e3, "bad": int((res.status != 0).sum().item()),
                          "sum_check": float(res.sum_wait.sum().item())}), flush=True)
        return
    for name, model, servers, arr, srv, trials, nobj in (("M/M/1", cb.MODEL_MM1, 1, 1 / 0.9, 1.0, 65536, 100000),
                                                         ("M/M/c c=8", cb.MODEL_MMC, 8, 1 / 6.4, 1.0, 32768, 100000)):
        am = torch.full((trials,), arr, dtype=torch.float64, device=dev)
        sm = torch.full((trials,), srv, dtype=torch.float64, device=dev)
        out = {}
        for label, variant in (("fast", 0), ("general", cb.VARIANT_GENERAL)) + ((("static", cb.VARIANT_STATIC), ("producer_consumer", 2)) if model == cb.MODEL_MM1 else ()):
            bufs = cb.TrialBuffers(trials, dev, 0, model, servers, variant)
            cb.launch_trials(am[:256], sm[:256], num_objects=1000, master_seed=1, model=model, servers=servers, variant=variant)
            res, ms = timed(lambda: cb.launch_trials(am, sm, num_objects=nobj, master_seed=MASTER, model=model, servers=servers,
                                                     variant=variant, buffers=bufs))
            ev = int(res.events.sum().item())
            out[label] = {"ms": ms, "events_per_s": ev / ms * 1e3, "events": ev, "bad": int((res.status != 0).sum().item()),
                          "sum_check": float(res.sum_wait.sum().item())}
        row = {"model": name, "trials": trials, "objects": nobj, **out,
               "general_over_fast_time": out["general"]["ms"] / out["fast"]["ms"],
               "static_over_fast_time": out["static"]["ms"] / out["fast"]["ms"] if "static" in out else None,
               "same_answers": all(o["sum_check"] == out["fast"]["sum_check"] and o["events"] == out["fast"]["events"] for o in out.values())}
        rows.append(row)
        print(json.dumps(row), flush=True)
    for trials, customers, T in ((4096, 1000, 50), (16384, 1000, 20)):
        am = torch.full((trials,), 4.0, dtype=torch.float64, device=dev)
        sm = torch.full((trials,), 1.0, dtype=torch.float64, device=dev)
        bufs = cb.TrialBuffers(trials, dev, 0, cb.MODEL_RENEGE, customers, 0)
        res, ms = timed(lambda: cb.launch_trials(am, sm, num_objects=T, master_seed=MASTER, model=cb.MODEL_RENEGE, servers=customers,
                                                 params=[0.5], buffers=bufs))
        ev = int(res.events.sum().item())
        row = {"model": "reneging customers", "trials": trials, "processes_per_trial": customers, "sim_time": T, "ms": ms,
               "events_per_s": ev / ms * 1e3, "events": ev, "bad": int((res.status != 0).sum().item()),
               "workspace_MB": bufs.workspace_bytes / 1e6}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if a.out:
        Path(a.out).parent.mkdir(parents=True, exist_ok=True)
        Path(a.out).write_text(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
