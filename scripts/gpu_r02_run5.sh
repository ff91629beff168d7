#!/bin/bash
# round 2, GPU call 5: general engine third form, AWACS 4096 x 24 h on the tutorial's own 14.4 GB map, mm1 A/B, the bench line
mkdir -p gpurun_out
timeout 600 python scripts/engine_bench.py --out gpurun_out/r02_engine_bench_v3.json > gpurun_out/r02_run5_engine_bench.log 2>&1
tail -5 gpurun_out/r02_run5_engine_bench.log | cut -c1-700
timeout 300 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e > gpurun_out/r02_run5_bench_shipped.json 2>/dev/null
CIMBA_B200_LIB=$PWD/cimba_b200/lib/variants/mm1_refill_branch.so timeout 300 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --no-e2e > gpurun_out/r02_run5_bench_refill_branch.json 2>/dev/null
python - <<'PY'
import json
for f in ("r02_run5_bench_shipped", "r02_run5_bench_refill_branch"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"])
PY
timeout 1500 python scripts/awacs_full.py --width 1000 --height 1000 --hours 24 --trials 4096 --out gpurun_out/r02_awacs_fullmap_24h.json > gpurun_out/r02_run5_awacs_full_24h.log 2>&1; echo "awacs full 24h rc=$?"
tail -1 gpurun_out/r02_run5_awacs_full_24h.log | cut -c1-900
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_run5_bench_full.json 2> gpurun_out/r02_run5_bench_full.err; echo "bench rc=$?"
head -c 1500 gpurun_out/r02_run5_bench_full.json
