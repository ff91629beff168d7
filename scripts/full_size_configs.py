"""BASELINE.json configs 3 and 4 at their full per-GPU size (one B200): invariants + throughput."""
import json, sys, time
sys.path.insert(0, ".")
import torch
import cimba_b200 as cb
K = 0x34F05C64D7AD598F
for name, model, n, arr, srv, servers in (("config 3 shard: M/M/c c=8, 32768 of 262144 replications x 1e6 customers", cb.MODEL_MMC, 32768, 1 / 6.4, 1.0, 8),
                                          ("config 4: G/G/1, 1048576 replications x 1e6 objects", cb.MODEL_GG1, 1048576, 1.25, 1.0, 1)):
    cb.run_trials(1024, arr_mean=arr, srv_mean=srv, num_objects=1000, master_seed=K, model=model, servers=servers)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = cb.run_trials(n, arr_mean=arr, srv_mean=srv, num_objects=1_000_000, master_seed=K, model=model, servers=servers)
    dt = time.perf_counter() - t0
    ev = r.total_events()
    avg = (r.sum_wait / r.objects.double())
    s = cb.DataSummary.from_list(cb.summarize_on_device(r.sum_wait, r.objects).cpu().tolist())
    print(json.dumps({"config": name, "trials": n, "events": ev, "seconds": dt, "events_per_s": ev / dt,
                      "failed_trials": int((r.status != 0).sum()), "all_objects_served": bool((r.objects == 1_000_000).all()),
                      "mean_time_in_system": s.mean(), "ci95_half_width": s.half_width_95()}), flush=True)
