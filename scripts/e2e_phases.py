"""Phase times of the host-buffer path (CIMBA_B200_TIMING=1) at the bench size."""
import os, sys, time
sys.path.insert(0, ".")
os.environ["CIMBA_B200_TIMING"] = "1"
import numpy as np, torch
import cimba_b200 as cb
T = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
NOBJ = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
exp = np.zeros(T, dtype=cb.TRIAL_DTYPE)
exp["arr_mean"], exp["srv_mean"] = 1 / 0.9, 1.0
cb.cimba_run_experiment(exp, num_objects=1000, master_seed=1)
for _ in range(2):
    t0 = time.perf_counter()
    cb.cimba_run_experiment(exp, num_objects=NOBJ, master_seed=1)
    print("wall %.1f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
# device-resident for comparison
res = cb.run_trials(T, arr_mean=1 / 0.9, srv_mean=1.0, num_objects=NOBJ, master_seed=1)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = cb.run_trials(T, arr_mean=1 / 0.9, srv_mean=1.0, num_objects=NOBJ, master_seed=1)
torch.cuda.synchronize()
print("device-resident wall %.1f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
