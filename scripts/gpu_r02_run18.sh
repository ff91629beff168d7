#!/bin/bash
# round 2, GPU call 18: harbor fused vs general engine at 4096 and 65536 trials; smoke(); memcheck of the static-tier launches
mkdir -p gpurun_out
timeout 900 python scripts/coverage_bench.py --harbor-only --out gpurun_out/r02_harbor_bench.json > gpurun_out/r02_run18_harbor.log 2>&1; tail -3 gpurun_out/r02_run18_harbor.log | cut -c1-600
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/r02_run18_memcheck_static.log 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
import cimba_b200 as cb
K = 0x34F05C64D7AD598F
for kw in (dict(model=cb.MODEL_MM1, n=100, arr=1 / 0.9, srv=1.0, size=300), dict(model=cb.MODEL_MM1, n=70, arr=0.5, srv=1.0, size=3000),
           dict(model=cb.MODEL_GG1, n=100, arr=1.25, srv=1.0, size=300), dict(model=cb.MODEL_GG1, n=40, arr=0.6, srv=1.0, size=3000)):
    r = cb.run_trials(kw["n"], arr_mean=kw["arr"], srv_mean=kw["srv"], num_objects=kw["size"], master_seed=K, model=kw["model"], variant=cb.VARIANT_STATIC)
    assert int(r.status.abs().sum()) == 0
    print(kw["model"], r.total_events(), flush=True)
mid = cb.load_model("cimba_b200/lib/models/libtandem_static_user_model.so")
r = cb.run_trials(64, arr_mean=1.05, srv_mean=1.0, num_objects=3000, master_seed=K, model=mid, servers=1)
print("tandem", r.total_events(), int(r.status.abs().sum()))
PY
echo "memcheck rc=$?"; tail -4 gpurun_out/r02_run18_memcheck_static.log
