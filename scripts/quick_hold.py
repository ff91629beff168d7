"""Throughput of the hold model: deep levels in HBM/L2 (variant 0) vs all in shared memory (variant 1)."""
import sys, time
import torch
sys.path.insert(0, ".")
import cimba_b200 as cb
cb.run_trials(64, arr_mean=1.0, srv_mean=1.0, num_objects=2, master_seed=1, model=cb.MODEL_HOLD, servers=1000)
for v in (1, 2, 3, 4):
    cb.run_trials(64, arr_mean=1.0, srv_mean=1.0, num_objects=2, master_seed=1, model=cb.MODEL_HOLD, servers=1000, variant=v)
CASES = [(v, 4096, 1000, 100) for v in (1, 2, 3, 4)] + [(v, 16384, 1000, 50) for v in (2, 3, 4)] + \
        [(v, 65536, 1000, 20) for v in (2, 3, 4)] + [(v, 4096, 100, 500) for v in (1, 2, 3, 4)] + \
        [(v, 4096, 10000, 10) for v in (2, 3, 4)] + [(v, 2048, 33000, 5) for v in (2, 3, 4)]
for variant, n, workers, dur in CASES:
    torch.cuda.synchronize(); t0 = time.time()
    r = cb.run_trials(n, arr_mean=1.0, srv_mean=1.0, num_objects=dur, master_seed=1, model=cb.MODEL_HOLD, servers=workers, variant=variant)
    dt = time.time() - t0
    print("HOLD variant %d: %d trials x %d workers x %d: %.3f Gev/s (%.3f s)" % (variant, n, workers, dur, r.total_events() / dt / 1e9, dt), flush=True)
