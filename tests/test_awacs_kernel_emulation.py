"""CPU test: the SOURCE TEXT of awacs_kernel (cimba_b200/csrc/awacs_model.cuh) executed on the CPU - 32 lanes as
coroutines, warp primitives as rendezvous (tests/awacs_kernel_emulation.cpp) - against the plain-C oracle.

It checks the kernel's logic (event selection, the target state machine, the three passes of a radar tick, the order
of the random draws) where no GPU is present and for simulated times that would be slow to compare on a GPU box; the
device's arithmetic and generator are covered by the GPU tests.  Longer runs done by hand with the same binary are
recorded in profiles/r01_awacs.md (1200 s and 7200 s: identical to the oracle in every pop, position and state)."""
import json
import subprocess
from pathlib import Path

from oracle_libs import load_port

ROOT = Path(__file__).resolve().parents[1]


def test_kernel_source_emulated_on_the_cpu_matches_the_oracle(tmp_path):
    load_port()                                         # builds oracle/liboracle_port.so if needed
    exe = tmp_path / "awacs_kernel_emulation"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wno-unknown-pragmas",
                    str(ROOT / "tests/awacs_kernel_emulation.cpp"), "-o", str(exe),
                    f"-L{ROOT / 'oracle'}", "-loracle_port", f"-Wl,-rpath,{ROOT / 'oracle'}", "-lm"],
                   check=True, capture_output=True)
    for seconds, trial in ((300, 3), (90, 17)):
        out = json.loads(subprocess.run([str(exe), "6", "5", str(seconds), str(trial)], check=True, capture_output=True,
                                        text=True, timeout=600).stdout)
        assert out["events"][0] == out["events"][1] and out["found"][0] == out["found"][1] and out["t_end_equal"]
        assert out["status"] == 0 and out["first_trace_diff"] == -1 and out["compared_pops"] == out["events"][0]
        assert (out["position_diffs"], out["tds_diffs"], out["mode_diffs"], out["found_diffs"]) == (0, 0, 0, 0), out
