import sys, time
import torch
sys.path.insert(0, ".")
import cimba_b200 as cb
cb.run_trials(1024, arr_mean=1 / 0.9, srv_mean=1.0, num_objects=100, master_seed=1)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    r = cb.run_trials(65536, arr_mean=1 / 0.9, srv_mean=1.0, num_objects=200000, master_seed=1)
    dt = time.time() - t0
    print("  M/M/1 65536 x 200000: %.4g ev/s" % (r.total_events() / dt), flush=True)
