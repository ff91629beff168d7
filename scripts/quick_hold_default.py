import sys, time
import torch
sys.path.insert(0, ".")
import cimba_b200 as cb
cb.run_trials(64, arr_mean=1.0, srv_mean=1.0, num_objects=2, master_seed=1, model=cb.MODEL_HOLD, servers=1000)
for n, workers, dur in ((4096, 1000, 100), (16384, 1000, 50), (4096, 100, 500), (4096, 10000, 10)):
    torch.cuda.synchronize(); t0 = time.time()
    r = cb.run_trials(n, arr_mean=1.0, srv_mean=1.0, num_objects=dur, master_seed=1, model=cb.MODEL_HOLD, servers=workers)
    dt = time.time() - t0
    print("HOLD default: %d trials x %d workers x %d: %.3f Gev/s (%.3f s)" % (n, workers, dur, r.total_events() / dt / 1e9, dt), flush=True)
