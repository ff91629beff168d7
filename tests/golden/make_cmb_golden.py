"""Golden vectors for the general engine's models, produced by the UNMODIFIED reference (oracle/_ref/librefdrv.so: the
same models written against the reference's API in oracle/ref_build/ref_driver.c).

    python tests/golden/make_cmb_golden.py      -> tests/golden/cmb_engine_vectors.json

Each case: model, servers, num_objects, arr_mean, srv_mean, params; per trial (seed cmb_random_fmix64(MASTER, i)):
events, objects, t_end and sum_wait as hex floats, the first four counters, and the SHA-256 of the first `trace`
pops (key, time).  Used by tests/test_cmb_engine.py (the engine's source text run on the CPU) and
tests/test_gpu_cmb_engine.py (the same on the device) wherever the live reference build is absent."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
from oracle_libs import load_ref, run_trials, trace_trial      # noqa: E402
import ctypes as C                                              # noqa: E402

MASTER = 0x34F05C64D7AD598F
TRACE = 2000
CASES = [
    # model, servers, num_objects, arr_mean, srv_mean, params, trials
    (0, 1, 20000, 1 / 0.9, 1.0, [], 6),
    (0, 1, 60000, 1 / 0.99, 1.0, [], 4),          # heavy traffic: the queue passes the fast kernel's window + ring
    (0, 1, 20000, 1 / 1.05, 1.0, [], 4),          # overload: the queue grows for the whole trial
    (1, 1, 20000, 1.25, 1.0, [], 6),
    (1, 1, 40000, 1.02, 1.0, [], 4),
    (2, 8, 20000, 1 / 6.4, 1.0, [], 6),
    (2, 3, 20000, 1 / 2.9, 1.0, [], 4),
    (2, 64, 20000, 1 / 60.0, 1.0, [], 4),         # more servers than the fast kernel's event list holds
    (2, 8, 20000, 1 / 8.4, 1.0, [], 4),           # overload: the wait list grows
    (7, 50, 40, 1.0, 1.0, [], 4),                 # the hold model: 50 and 2000 workers + ticker + end event
    (7, 2000, 6, 1.0, 1.0, [], 3),
    (10, 10, 1500, 2.0, 8.0, [], 4),              # the harbor (test/test_condition.c): 10 tugs; 4 tugs under load
    (10, 4, 1200, 1.5, 8.0, [], 3),
    (16, 40, 300, 3.0, 1.0, [0.7], 4),
    (16, 1000, 60, 4.0, 1.0, [0.5], 4),
    (16, 1500, 40, 2.0, 1.0, [1.5], 3),
    (17, 4, 20000, 1.3, 1.0, [], 4),
    (18, 20, 100, 1.0, 1.0, [], 6),               # test/test_resourcepool.c as it stands (20 units, 100 time units)
    (18, 7, 300, 1.0, 1.0, [], 4),
    (17, 1, 20000, 1.05, 1.0, [], 4),
    (3, 10, 3000, 1.0, 1.0, [], 4),               # test/test_objectqueue.c's workload; a tight queue (both guards busy)
    (3, 2, 2000, 0.6, 1.0, [], 4),
    (11, 10, 20000, 1.0, 1.0, [], 3),             # ... with the length history on
    (13, 10, 20000, 1.0, 1.0, [], 3),             # test/test_priorityqueue.c
    (13, 3, 3000, 0.7, 1.0, [], 4),
    (5, 10, 3000, 1.0, 1.0, [], 4),               # buffer + resource with pre-emption under a nuisance
    (5, 4, 2000, 0.8, 1.0, [], 4),
    (12, 10, 10000, 1.0, 1.0, [], 3),             # test/test_buffer.c as it stands (trial 0 of the reference's seed is in test_cmb_engine.py)
    (14, 1, 25, 1.0, 1.0, [], 6),                 # test/test_resource.c as it stands
    (14, 1, 5000, 1.0, 1.0, [], 3),
    (4, 8, 3000, 1.0, 1.0, [], 4),                # pool with pre-emption, priority changes, interrupts
    (4, 3, 2000, 1.0, 1.0, [], 4),
    (6, 6, 3000, 0.7, 1.0, [], 4),                # priority queue by handle + condition with two predicates
    (6, 2, 2000, 0.5, 1.0, [], 4),
    (8, 1, 3000, 1.0, 0.8, [], 4),                # timers, yield / resume, waits on processes and events, an observing guard
    (8, 1, 2000, 0.4, 1.5, [], 4),
    (9, 1, 20000, 1 / 0.9, 1.0, [], 6),           # tutorial 1 / test/test_cimba.c: M/M/1 with the queue's history on
    (9, 1, 30000, 1 / 0.97, 1.0, [], 3),
    (19, 1, 20000, 1 / 0.9, 1.0, [1000.0], 6),    # tutorial/tut_1_7.c's trial: warm-up 1000, duration 20000, rho 0.9
    (19, 1, 30000, 1 / 0.5, 1.0, [0.0], 4),       # no warm-up, rho 0.5
    (19, 1, 5000, 1 / 0.975, 1.0, [250.5], 4),    # the tutorial's heaviest load
]


def main():
    ref = load_ref()
    assert ref is not None, "oracle/_ref/librefdrv.so is not built (make -C oracle ref)"
    ref.ref_set_param.argtypes = [C.c_int, C.c_double]
    out = {"master": MASTER, "trace": TRACE, "cases": []}
    for model, servers, nobj, arr, srv, params, n in CASES:
        ref.ref_set_param(0, params[0] if params else 0.0)
        res = run_trials(ref, "ref", model, servers, MASTER, 0, n, nobj, arr, srv, par=0)
        trials = []
        for i, r in enumerate(res):
            _, keys, times = trace_trial(ref, "ref", model, servers, ref.ref_fmix64(MASTER, i), nobj, arr, srv, TRACE)
            h = hashlib.sha256(np.array(keys, dtype=np.uint64).tobytes() + np.array(times, dtype=np.float64).tobytes())
            trials.append({"events": r.events, "objects": r.objects, "t_end": float(r.t_end).hex(), "sum_wait": float(r.sum_wait).hex(),
                           "counters": list(r.counter)[:4], "counters8": list(r.counter), "all8": int(model in (3, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14, 18, 19)), "max_queue": r.max_queue, "max_fel": r.max_fel, "pops": len(keys), "trace_sha256": h.hexdigest()})
        out["cases"].append({"model": model, "servers": servers, "num_objects": nobj, "arr_mean": float(arr).hex(),
                             "srv_mean": float(srv).hex(), "params": params, "trials": trials})
        print(model, servers, nobj, [t["events"] for t in trials])
    ref.ref_set_param(0, 0.0)
    (ROOT / "tests/golden/cmb_engine_vectors.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
