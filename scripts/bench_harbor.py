"""Throughput of the harbor kernel (BASELINE config 5's shape: 4096 replications)."""
import sys, time
sys.path.insert(0, ".")
import torch
import cimba_b200 as cb
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
H = int(sys.argv[2]) if len(sys.argv) > 2 else 8736
for variant in (0, 1, 2):
    cb.run_trials(T, arr_mean=2.0, srv_mean=8.0, num_objects=100, master_seed=1, model=cb.MODEL_HARBOR, servers=10, variant=variant)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = cb.run_trials(T, arr_mean=2.0, srv_mean=8.0, num_objects=H, master_seed=1, model=cb.MODEL_HARBOR, servers=10, variant=variant)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ev = int(res.events.sum())
    print(f"harbor variant={variant} trials={T} hours={H} events={ev} time={dt:.3f}s events/s={ev/dt:.4g} bad={int((res.status!=0).sum())}", flush=True)
