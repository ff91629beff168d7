"""Calibrate "warp-instructions per event-loop iteration" of the fast kernels against ncu, for bench.py's roofline.

On the GPU box (one call):
    ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum \
        --clock-control none -k regex:'mm1_kernel|gg1_kernel|pool_fast_kernel' --csv --log-file gpurun_out/issue_ncu.csv \
        python scripts/calibrate_issue.py --run gpurun_out/issue_diag.json
Here (no GPU):
    python scripts/calibrate_issue.py --combine gpurun_out/issue_diag.json gpurun_out/issue_ncu.csv   -> profiles/issue_calibration.json

--run launches each kernel ONCE at the benchmark's trial count (objects scaled down: the loop's instruction count does
not depend on the trial length) with job.diag set, and writes the kernels' own iteration counts; ncu's
smsp__inst_executed.sum of the same launch divided by that count is the calibration."""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
MASTER = 0x34F05C64D7AD598F
KERNELS = [("mm1_kernel", 0, 1, 1 / 0.9, 1.0, 65536), ("gg1_kernel", 1, 1, 1.25, 1.0, 65536), ("pool_fast_kernel", 2, 8, 1 / 6.4, 1.0, 32768)]


def run(out):
    import torch
    import cimba_b200 as cb
    dev = torch.device("cuda", 0)
    rows = {}
    for name, model, servers, arr, srv, trials in KERNELS:
        am = torch.full((trials,), arr, dtype=torch.float64, device=dev)
        sm = torch.full((trials,), srv, dtype=torch.float64, device=dev)
        diag = torch.zeros(4, dtype=torch.int64, device=dev)
        res = cb.launch_trials(am, sm, num_objects=100_000, master_seed=MASTER, model=model, servers=servers, diag=diag)
        torch.cuda.synchronize()
        d = diag.cpu().tolist()
        rows[name] = {"iterations": d[0], "warps": d[1], "events": int(res.events.sum().item()), "trials": trials}
    Path(out).write_text(json.dumps(rows, indent=1))


def combine(diag_file, ncu_csv):
    diag = json.loads(Path(diag_file).read_text())
    per = {}
    with open(ncu_csv) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        k = next((n for n in diag if n in r["Kernel Name"]), None)
        if k:
            per.setdefault(k, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    out = {}
    for k, m in per.items():
        inst = m["smsp__inst_executed.sum"]
        out[k] = {"warp_instructions_per_iteration": inst / diag[k]["iterations"],
                  "warp_instructions_per_event": inst / diag[k]["events"],
                  "issue_active_pct_under_ncu": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "alu_pipe_pct_under_ncu": m.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                  "active_threads_per_instruction": m.get("smsp__thread_inst_executed_per_inst_executed.ratio"),
                  "source": f"ncu smsp__inst_executed.sum = {inst:.0f} over {diag[k]['iterations']} loop iterations the kernel counted "
                            f"itself ({diag[k]['trials']} trials x 1e5 objects), scripts/calibrate_issue.py"}
    (ROOT / "profiles/issue_calibration.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        run(sys.argv[2])
    elif len(sys.argv) >= 4 and sys.argv[1] == "--combine":
        combine(sys.argv[2], sys.argv[3])
    else:
        sys.exit(__doc__)
