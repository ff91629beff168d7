"""Golden vectors of tutorial/tut_2_1.c from the UNMODIFIED tutorial source run as a program, one fresh process per trial
(oracle/ref_build/tut2_main.c -> oracle/_ref/tut2_ref; why a process per trial: see that file).

    python tests/golden/make_tutorial2_golden.py      -> tests/golden/tutorial2_vectors.json

Per trial (seed = cmb_random_fmix64(MASTER, i)): events executed, final clock, and the next raw output of the random stream
after the run (a fingerprint of every draw the trial made)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
from oracle_libs import load_ref      # noqa: E402

MASTER = 0x34F05C64D7AD598F
TRIALS = 32


def main():
    ref = load_ref()
    assert ref is not None
    rows = []
    for i in range(TRIALS):
        seed = ref.ref_fmix64(MASTER, i)
        ev, te, nxt = subprocess.run([str(ROOT / "oracle/_ref/tut2_ref"), hex(seed)], capture_output=True, text=True, check=True).stdout.split()
        rows.append({"events": int(ev), "t_end": float.fromhex(te).hex(), "next_raw": int(nxt)})
    (ROOT / "tests/golden/tutorial2_vectors.json").write_text(json.dumps({"master": MASTER, "trials": rows}, indent=1))
    print(TRIALS, "trials,", sum(r["events"] for r in rows), "events")


if __name__ == "__main__":
    main()
