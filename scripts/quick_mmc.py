"""M/M/c throughput (BASELINE config 3: c = 8, rho = 0.8)."""
import sys, time
import torch
sys.path.insert(0, ".")
import cimba_b200 as cb
cb.run_trials(1024, arr_mean=1 / 6.4, srv_mean=1.0, num_objects=100, master_seed=1, model=cb.MODEL_MMC, servers=8)
cb.run_trials(1024, arr_mean=1 / 6.4, srv_mean=1.0, num_objects=100, master_seed=1, model=cb.MODEL_MMC, servers=8, variant=1)
for variant in (0, 1):
    for n, nobj in ((32768, 50000), (262144, 20000)):
        torch.cuda.synchronize(); t0 = time.time()
        r = cb.run_trials(n, arr_mean=1 / 6.4, srv_mean=1.0, num_objects=nobj, master_seed=1, model=cb.MODEL_MMC, servers=8, variant=variant)
        dt = time.time() - t0
        print("M/M/c c=8 variant %d: %d trials x %d customers: %.4g ev/s (%.3f s) bad=%d" % (variant, n, nobj, r.total_events() / dt, dt, int((r.status != 0).sum())), flush=True)
