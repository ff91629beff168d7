import sys
sys.path.insert(0, ".")
import cimba_b200 as cb
for w in (40, 600):
    r = cb.run_trials(9, arr_mean=1.0, srv_mean=1.0, num_objects=4, master_seed=7, model=cb.MODEL_HOLD, servers=w, variant=1)
    print(r.total_events())
