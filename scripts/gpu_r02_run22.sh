#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cmb_engine.py -q -x 2>&1 | tail -4
timeout 600 python - <<'PY'
import sys, json
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
import torch, cimba_b200 as cb
from engine_bench import timed, MASTER
dev = torch.device("cuda", 0)
# tutorial/tut_1_7.c at its own size: 39 utilisations x 10 replications would be 390 trials; a GPU wants more - 39 x 1680 = 65 520
rhos = [0.025 * (k + 1) for k in range(39)]
reps = 1680
n = len(rhos) * reps
am = torch.tensor([1.0 / rhos[i // reps] for i in range(n)], dtype=torch.float64, device=dev)
sm = torch.ones(n, dtype=torch.float64, device=dev)
for label, variant in (("static", 0), ("general", cb.VARIANT_GENERAL)):
    dur = 100000 if variant == 0 else 10000
    bufs = cb.TrialBuffers(n, dev, 0, cb.MODEL_TUTORIAL1, 1, variant)
    cb.launch_trials(am[:256], sm[:256], num_objects=100, master_seed=1, model=cb.MODEL_TUTORIAL1, variant=variant, params=[10.0])
    res, ms = timed(lambda: cb.launch_trials(am, sm, num_objects=dur, master_seed=MASTER, model=cb.MODEL_TUTORIAL1, variant=variant, params=[1000.0], buffers=bufs))
    ev = int(res.events.sum().item())
    print(json.dumps({"tutorial1": label, "trials": n, "duration": dur, "ms": ms, "events_per_s": ev / ms * 1e3, "events": ev, "bad": int((res.status != 0).sum().item())}), flush=True)
PY
