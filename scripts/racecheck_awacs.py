"""A small MODEL_AWACS launch for compute-sanitizer --tool racecheck / synccheck / initcheck: the TMA-fed radar pass (mbarrier
pipeline, shared-memory tiles) and the warp-cooperative event list are the only kernels of the library with inter-lane
shared-memory hand-offs."""
import sys
sys.path.insert(0, ".")
import torch
import cimba_b200 as cb
K = 0x34F05C64D7AD598F
cols, rows = 300, 200
yy, xx = torch.meshgrid(torch.arange(rows, dtype=torch.float32), torch.arange(cols, dtype=torch.float32), indexing="ij")
ridge = (400.0 + 300.0 * torch.sin(xx / 17.0) * torch.cos(yy / 11.0)).clamp_min(0.0).reshape(-1).cuda()
cb.awacs_set_terrain(ridge, cols, rows, (27.0, 31.0, -27.0 * (cols - 1) / 2, 27.0 * (cols - 1) / 2,
                                         -31.0 * (rows - 1) / 2, 31.0 * (rows - 1) / 2))
res, per = cb.awacs_run(int(sys.argv[1]) if len(sys.argv) > 1 else 6, duration_s=int(sys.argv[2]) if len(sys.argv) > 2 else 6, master_seed=K)
torch.cuda.synchronize()
print("awacs", res.total_events(), int(res.objects.sum()), flush=True)
