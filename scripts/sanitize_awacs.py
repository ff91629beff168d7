"""One small MODEL_AWACS launch for compute-sanitizer (synthetic terrain; no oracle involved)."""
import sys
sys.path.insert(0, ".")
import torch
import cimba_b200 as cb
cols, rows = 300, 200
yy, xx = torch.meshgrid(torch.arange(rows, dtype=torch.float32), torch.arange(cols, dtype=torch.float32), indexing="ij")
ridge = (400.0 + 300.0 * torch.sin(xx / 17.0) * torch.cos(yy / 11.0)).clamp_min(0.0).reshape(-1).cuda()
cb.awacs_set_terrain(ridge, cols, rows, (27.0, 31.0, -27.0 * (cols - 1) / 2, 27.0 * (cols - 1) / 2,
                                         -31.0 * (rows - 1) / 2, 31.0 * (rows - 1) / 2))
res, per = cb.awacs_run(5, duration_s=int(sys.argv[1]) if len(sys.argv) > 1 else 8, master_seed=0x34F05C64D7AD598F, trace_cap=64)
print("awacs events", res.total_events(), "found", int(res.objects.sum()), "bad", int((res.status != 0).sum()), flush=True)
