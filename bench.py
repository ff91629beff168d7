#!/usr/bin/env python3
"""bench.py - M/M/1 replication benchmark (BASELINE.json metric) on N B200s.

One "step" = one pass of the hot path over one batch: every rank runs
`--trials` independent M/M/1 replications of `--objects` customers each
(benchmark/MM1_multi.c with NUM_TRIALS=65536, rho=0.9) in ONE launch of the
persistent simulation kernel.  Weak scaling: per-GPU work is fixed, rank r runs
global trial indices [r*T, (r+1)*T); no data-path collective - the only exchange
is the 8-double cmb_datasummary all-gather after the timed region.

  value  = FEL pops ("events" = cmb_event_execute_next() calls) per second,
           all ranks, device-resident inputs, CUDA-event timed, max over ranks.
  e2e    = same metric through the host-buffer C-ABI entry
           (cimba_b200_run_experiment: pinned staging, H2D, kernel, D2H inside).
  roofline.achieved = events/s * 72 B/event (packed-record algorithmic bytes,
           SURVEY.md section 8d) against the measured HBM copy peak.  The path is
           NOT HBM-bound by design (per-trial state lives on chip) - see DESIGN.md.
  cpu_baseline / --impl reference = the reference's own pthread executive
           (oracle/_ref, unmodified sources) on this box's host cores, on a
           bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

MASTER_SEED = 0x34F05C64D7AD598F
ARRIVAL_RATE, SERVICE_RATE = 0.9, 1.0          # benchmark/MM1_multi.c:27-28
BYTES_PER_EVENT = 72.0                         # SURVEY.md section 8d (packed layout)
BYTES_PER_EVENT_REFLAYOUT = 150.9              # same, reference record layout


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="native", choices=["native", "reference"])
    p.add_argument("--trials", type=int, default=65536, help="replications per GPU per step")
    p.add_argument("--objects", type=int, default=1_000_000, help="customers per replication")
    p.add_argument("--mapping", type=int, default=1, choices=[1, 32], help="1 lane/trial or 32 (warp/trial)")
    p.add_argument("--variant", type=int, default=0, help="0 default kernel, 1 unfused formulation (A/B)")
    p.add_argument("--ref-trials", type=int, default=0, help="CPU sample size (0 = 16 per core)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


def measured_traffic():
    """DRAM bytes per launch of the timed kernel from the committed ncu capture
    (profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum); None if absent."""
    f = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(f.read_text())
    except Exception:
        return None


def measured_peak_gbs():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(trials: int, objects: int, first: int = 0):
    """Time the reference's own CPU path (cimba_run_experiment over all cores,
    oracle/_ref) on `trials` replications; falls back to the oracle port."""
    from oracle_libs import load_port, load_ref, run_trials
    ref = load_ref()
    cores = os.cpu_count() or 1
    if ref is not None:
        cores = ref.ref_cpu_cores()
        t0 = time.perf_counter()
        res = run_trials(ref, "ref", 0, 1, MASTER_SEED, first, trials, objects,
                         1.0 / ARRIVAL_RATE, 1.0 / SERVICE_RATE, par=1)
        dt = time.perf_counter() - t0
        kind = "reference"
    else:
        port = load_port()
        t0 = time.perf_counter()
        res = run_trials(port, "port", 0, 1, MASTER_SEED, first, trials, objects,
                         1.0 / ARRIVAL_RATE, 1.0 / SERVICE_RATE, par=cores)
        dt = time.perf_counter() - t0
        kind = "port"
    events = sum(r.events for r in res)
    return {"events": events, "seconds": dt, "cores": cores, "kind": kind, "results": res}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    trials = args.ref_trials or 16 * cores
    for _ in range(args.warmup):
        cpu_reference_run(max(cores, 1), min(args.objects, 100_000))
    t_total, ev_total, kind, ncores = 0.0, 0, "reference", cores
    for k in range(args.steps):
        r = cpu_reference_run(trials, args.objects, first=k * trials)
        t_total += r["seconds"]
        ev_total += r["events"]
        kind, ncores = r["kind"], r["cores"]
    value = ev_total / t_total
    sample = f"{trials} of {args.trials} replications x {args.objects} objects per step"
    line = {
        "impl": "reference", "metric": "M/M/1 simulated events/sec (FEL pops/s)", "value": value,
        "unit": "events/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "benchmark/MM1_multi (rho=0.9, 1e6 objects/trial, 65536 trials/GPU)",
                   "sample": sample, "seeding": "cmb_random_fmix64(master, trial)"},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": ncores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "events_convention_4_per_object": 4.0 * trials * args.objects * args.steps / t_total,
    }
    emit(line)


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    # The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under
    # NCCL_DEBUG=VERSION, for one), so file descriptor 1 is pointed at stderr for the whole run and the
    # line goes to the saved descriptor at the end.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import cimba_b200 as cb
    from cimba_b200.experiment import TrialBuffers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU path exists)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    T, NOBJ = args.trials, args.objects
    first = rank * T
    arr = torch.full((T,), 1.0 / ARRIVAL_RATE, dtype=torch.float64, device=dev)
    srv = torch.full((T,), 1.0 / SERVICE_RATE, dtype=torch.float64, device=dev)
    bufs = TrialBuffers(T, dev)

    def step():
        return cb.launch_trials(arr, srv, num_objects=NOBJ, master_seed=MASTER_SEED, first_trial=first,
                                mapping=args.mapping, buffers=bufs, variant=args.variant)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()                                     # every rank watches its own GPU
    launches0 = cb.lib.cimba_b200_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    per_launch = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                  for _ in range(args.steps)]
    barrier()
    ev0.record()
    for k in range(args.steps):
        per_launch[k][0].record()
        res = step()
        per_launch[k][1].record()
    ev1.record()
    barrier()
    launches = cb.lib.cimba_b200_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = [a.elapsed_time(b) for a, b in per_launch]
    clocks = sampler.stop()

    events_rank = int(res.events.sum().item())
    bad = int((res.status != 0).sum().item())
    t = torch.tensor([ms_total, float(events_rank), float(bad)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, events_all, bad = float(tmax[0]), int(tsum[1]), int(tsum[2])
    else:
        events_all = events_rank
    per_rank = None
    if world > 1:
        # the slowest rank sets the number; record every rank's own time and clocks beside it
        mine = {"rank": rank, "ms_per_step": float(t[0]) / args.steps, "sm_mhz": clocks.get("sm_mhz"),
                "reasons": clocks.get("reasons")}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered
    ms_per_step = ms_total / args.steps
    value = events_all / (ms_per_step * 1e-3)

    # cross-GPU statistics merge (the path's only exchange; outside the timed region)
    local_summary = cb.summarize_on_device(res.sum_wait, res.objects)
    merged = cb.merge_across_ranks(local_summary)

    # ---- end to end through the host-buffer C-ABI (pinned staging + H2D + kernel + D2H)
    e2e = None
    if not args.no_e2e:
        exp = np.zeros(T, dtype=cb.TRIAL_DTYPE)
        exp["arr_mean"], exp["srv_mean"] = 1.0 / ARRIVAL_RATE, 1.0 / SERVICE_RATE
        n_e2e = max(1, min(args.steps, 2))
        cb.cimba_run_experiment(exp, num_objects=min(NOBJ, 1000), master_seed=MASTER_SEED,
                                first_trial=first, mapping=args.mapping, device=local)     # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            cb.cimba_run_experiment(exp, num_objects=NOBJ, master_seed=MASTER_SEED,
                                    first_trial=first, mapping=args.mapping, device=local)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        te = torch.tensor([dt, float(exp["events"].sum())], dtype=torch.float64, device=dev)
        if world > 1:
            tm = te.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ts = te.clone()
            dist.all_reduce(ts, op=dist.ReduceOp.SUM)
            dt, ev_e2e = float(tm[0]), float(ts[1])
        else:
            ev_e2e = float(te[1])
        e2e = {"value": ev_e2e * n_e2e / dt, "unit": "events/s", "steps": n_e2e,
               "h2d_bytes_per_step": 16 * T, "d2h_bytes_per_step": 40 * T,
               "api": "cimba_b200_run_experiment (host trial-struct array, results in place)",
               "bit_identical_to_device_path": bool(
                   np.array_equal(exp["events"], res.events.cpu().numpy().astype(np.uint64)))}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    kernel_s = sum(kernel_ms) / len(kernel_ms) * 1e-3
    ach = events_rank * BYTES_PER_EVENT / kernel_s / 1e9
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "mm1_kernel" if args.variant == 0 else "queue_kernel<0>",
                "kernel_ms": kernel_s * 1e3, "algorithmic_bytes_per_event": BYTES_PER_EVENT,
                "achieved_reference_layout_gbs": events_rank * BYTES_PER_EVENT_REFLAYOUT / kernel_s / 1e9,
                "note": "per-trial state is on-chip by design; actual HBM traffic is ~0, see DESIGN.md"}

    tr = measured_traffic()
    if tr and args.variant == 0 and args.mapping == 1:
        # scale the captured launch to this launch by its event count (traffic is parameters +
        # results + the ~3 % of queue entries that spill: all proportional to trials x objects)
        roofline["traffic"] = tr["dram_bytes_per_event"] * events_rank
        roofline["traffic_source"] = tr["source"]
    roofline["sm_issue"] = {"note": "binding limit (see DESIGN.md section 5)",
                            "warp_instructions_per_event_step": tr.get("warp_instructions_per_step") if tr else None,
                            "issue_slots_busy": tr.get("issue_slots_busy") if tr else None}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        trials = args.ref_trials or 16 * cores
        r = cpu_reference_run(trials, NOBJ)
        # parity on the very same trials, at full per-trial size
        gpu_ev = res.events[:trials].cpu().tolist()
        gpu_te = res.t_end[:trials].cpu().tolist()
        gpu_sw = res.sum_wait[:trials].cpu().tolist()
        same = all((gpu_ev[i], gpu_te[i], gpu_sw[i]) == (w.events, w.t_end, w.sum_wait)
                   for i, w in enumerate(r["results"]))
        cpu = {"value": r["events"] / r["seconds"], "unit": "events/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"{trials} of {T} replications x {NOBJ} objects (same seeds as GPU trials 0..{trials - 1})",
               "seconds": r["seconds"], "gpu_results_bit_identical_on_sample": bool(same)}

    line = {
        "metric": "M/M/1 simulated events/sec (FEL pops/s)", "value": value, "unit": "events/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "benchmark/MM1_multi (rho=0.9, 1e6 objects/trial, 65536 trials/GPU)",
                   "trials_per_gpu": T, "objects_per_trial": NOBJ, "rho": ARRIVAL_RATE / SERVICE_RATE,
                   "mapping": "lane-per-trial" if args.mapping == 1 else "warp-per-trial",
                   "seeding": "cmb_random_fmix64(0x34f05c64d7ad598f, global trial index)",
                   "l2": "no input re-use between steps: 1 MB of inputs, all state regenerated on chip"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "per_rank": per_rank, "roofline": roofline,
        "cpu_baseline": cpu,
        "events_per_step": events_all, "failed_trials": bad,
        "events_convention_4_per_object": 4.0 * T * world * NOBJ / (ms_per_step * 1e-3),
        "summary": {"n": merged.count(), "mean_time_in_system": merged.mean(),
                    "ci95_half_width": merged.half_width_95(), "expected": 1.0 / (SERVICE_RATE - ARRIVAL_RATE)},
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
