"""G/G/1 throughput: predicated kernel (variant 0) vs readable formulation (variant 1)."""
import sys, time
import torch
sys.path.insert(0, ".")
import cimba_b200 as cb
for v in (0, 1):
    cb.run_trials(1024, arr_mean=1.25, srv_mean=1.0, num_objects=100, master_seed=1, model=cb.MODEL_GG1, variant=v)
for variant, n, nobj in ((1, 65536, 200000), (0, 65536, 200000), (1, 1048576, 20000), (0, 1048576, 20000), (0, 606208, 50000)):
    torch.cuda.synchronize(); t0 = time.time()
    r = cb.run_trials(n, arr_mean=1.25, srv_mean=1.0, num_objects=nobj, master_seed=1, model=cb.MODEL_GG1, variant=variant)
    dt = time.time() - t0
    print("G/G/1 variant %d: %d trials x %d objects: %.4g ev/s (%.3f s) bad=%d" % (variant, n, nobj, r.total_events() / dt, dt, int((r.status != 0).sum())), flush=True)
