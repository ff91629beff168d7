import sys
sys.path.insert(0, ".")
import torch
import cimba_b200 as cb
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
r = cb.run_trials(n, arr_mean=1.0, srv_mean=1.0, num_objects=20, master_seed=1, model=cb.MODEL_HOLD, servers=1000, variant=variant)
print("events", r.total_events())
