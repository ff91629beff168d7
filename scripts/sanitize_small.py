"""Small launches of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, ".")
import torch
import cimba_b200 as cb
K = 0x34F05C64D7AD598F
runs = [
    dict(model=cb.MODEL_MM1, n=200, arr=1 / 0.9, srv=1.0, size=300),
    dict(model=cb.MODEL_MM1, n=100, arr=0.8, srv=1.0, size=400),                 # rho > 1: spill ring
    dict(model=cb.MODEL_MM1, n=64, arr=1 / 0.9, srv=1.0, size=200, variant=1),
    dict(model=cb.MODEL_MM1_RECORDED, n=64, arr=1 / 0.9, srv=1.0, size=200),
    dict(model=cb.MODEL_GG1, n=128, arr=1.25, srv=1.0, size=300),
    dict(model=cb.MODEL_GG1, n=64, arr=1.25, srv=1.0, size=200, variant=1),
    dict(model=cb.MODEL_MMC, n=128, arr=1 / 6.4, srv=1.0, size=300, servers=8),
    dict(model=cb.MODEL_GUARDED, n=64, arr=1.0, srv=1.0, size=60, servers=10),
    dict(model=cb.MODEL_PREEMPT, n=64, arr=1.0, srv=1.0, size=60, servers=20),
    dict(model=cb.MODEL_BUFFER, n=64, arr=1.0, srv=1.0, size=60, servers=10),
    dict(model=cb.MODEL_PRIOQ, n=64, arr=1.0, srv=1.0, size=60, servers=8),
    dict(model=cb.MODEL_TIMERS, n=64, arr=1.0, srv=0.6, size=60),
    dict(model=cb.MODEL_HARBOR, n=40, arr=2.0, srv=8.0, size=200, servers=10),
    dict(model=cb.MODEL_HARBOR, n=40, arr=1.5, srv=10.0, size=300, servers=5),     # on-chip overflow -> repair pass
    dict(model=cb.MODEL_HARBOR, n=70, arr=2.0, srv=8.0, size=100, servers=10, variant=2),
    # the static tier (cmb_static.cuh), with and without trials for its repair pass
    dict(model=cb.MODEL_MM1, n=100, arr=1 / 0.9, srv=1.0, size=300, variant=cb.VARIANT_STATIC),
    dict(model=cb.MODEL_MM1, n=70, arr=0.5, srv=1.0, size=3000, variant=cb.VARIANT_STATIC),
    dict(model=cb.MODEL_GG1, n=100, arr=1.25, srv=1.0, size=300, variant=cb.VARIANT_STATIC),
    # the general engine: every model written against the authoring surface
    dict(model=cb.MODEL_MM1, n=64, arr=1 / 0.9, srv=1.0, size=300, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_GG1, n=64, arr=1.25, srv=1.0, size=300, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_MMC, n=64, arr=1 / 60.0, srv=1.0, size=300, servers=64),
    dict(model=cb.MODEL_RENEGE, n=8, arr=3.0, srv=1.0, size=10, servers=1000, params=[0.7]),
    dict(model=cb.MODEL_POOL_RECORDED, n=64, arr=1.0, srv=1.0, size=60, servers=20),
    dict(model=cb.MODEL_HOLD, n=8, arr=1.0, srv=1.0, size=3, servers=2000, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_HARBOR, n=32, arr=2.0, srv=8.0, size=150, servers=10, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_GUARDED, n=64, arr=0.5, srv=1.0, size=60, servers=100),
    dict(model=cb.MODEL_GUARDED_RECORDED, n=64, arr=1.0, srv=1.0, size=60, servers=10, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_PRIOQ_RECORDED, n=64, arr=0.5, srv=1.0, size=60, servers=40),
    dict(model=cb.MODEL_PREEMPT, n=64, arr=1.0, srv=1.0, size=60, servers=20, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_BUFFER, n=64, arr=1.0, srv=1.0, size=60, servers=10, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_BUFFER_RECORDED, n=64, arr=1.0, srv=1.0, size=60, servers=10, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_PRIOQ, n=64, arr=0.7, srv=1.0, size=60, servers=8, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_TIMERS, n=64, arr=1.0, srv=0.6, size=60, variant=cb.VARIANT_GENERAL),
    dict(model=cb.MODEL_RESOURCE_RECORDED, n=64, arr=1.0, srv=1.0, size=60, variant=cb.VARIANT_GENERAL),
] + [dict(model=cb.MODEL_HOLD, n=9, arr=1.0, srv=1.0, size=4, servers=w, variant=v)
     for v in (1, 2, 3, 4) for w in (40, 600)] + [dict(model=cb.MODEL_HOLD, n=5, arr=1.0, srv=1.0, size=2, servers=3000, variant=v)
                                                   for v in (2, 3, 4)]
for r in runs:
    res = cb.run_trials(r["n"], arr_mean=r["arr"], srv_mean=r["srv"], num_objects=r["size"], master_seed=K,
                        model=r["model"], servers=r.get("servers", 1), variant=r.get("variant", 0), params=r.get("params", ()))
    assert int(res.status.abs().sum()) == 0 or r["model"] == cb.MODEL_MM1, (r, res.status.cpu().tolist())
    print(r["model"], r.get("variant", 0), res.total_events(), flush=True)
x = torch.rand(1000, dtype=torch.float64, device="cuda")
w = torch.rand(1000, dtype=torch.float64, device="cuda")
cb.summarize_weighted_on_device(x, w)
cb.rng_draws(K, 1, 1000, 1.0)
cb.rng_draws_ex(K, 17, 500, [1.0, 2.0, 6.0])
# MODEL_AWACS on a synthetic 300 x 200 ridge (the oracle's terrain generator is not needed for a memory check)
cols, rows = 300, 200
yy, xx = torch.meshgrid(torch.arange(rows, dtype=torch.float32), torch.arange(cols, dtype=torch.float32), indexing="ij")
ridge = (400.0 + 300.0 * torch.sin(xx / 17.0) * torch.cos(yy / 11.0)).clamp_min(0.0).reshape(-1).cuda()
cb.awacs_set_terrain(ridge, cols, rows, (27.0, 31.0, -27.0 * (cols - 1) / 2, 27.0 * (cols - 1) / 2,
                                         -31.0 * (rows - 1) / 2, 31.0 * (rows - 1) / 2))
res, per = cb.awacs_run(5, duration_s=12, master_seed=K, trace_cap=64)
print("awacs", res.total_events(), int(res.objects.sum()), flush=True)
torch.cuda.synchronize()
print("done")
