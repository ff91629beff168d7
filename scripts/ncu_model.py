import sys
sys.path.insert(0, ".")
import torch
import cimba_b200 as cb
model = int(sys.argv[1]); n = int(sys.argv[2]); size = int(sys.argv[3]); servers = int(sys.argv[4])
arr = float(sys.argv[5]); srv = float(sys.argv[6]); variant = int(sys.argv[7]) if len(sys.argv) > 7 else 0
r = cb.run_trials(n, arr_mean=arr, srv_mean=srv, num_objects=size, master_seed=1, model=model, servers=servers, variant=variant)
print("events", r.total_events(), "bad", int((r.status != 0).sum()))
