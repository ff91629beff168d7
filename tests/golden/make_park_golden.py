"""Golden vectors of the theme-park model (tutorial/tut_3_1.c) from the UNMODIFIED tutorial source, built by oracle/Makefile into
oracle/_ref/libtut3_ref.so with the trial's seed, its printing and its dispatcher loop redirected (oracle/ref_build/tut3_driver.c).

    python tests/golden/make_park_golden.py      -> tests/golden/park_vectors.json

Per trial (seed = cmb_random_fmix64(MASTER, i), as every experiment of this repository seeds its trials): events executed, final
clock, and the tutorial's five results - mean time in park, riding, waiting, walking, mean number of rides - as hex floats."""
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
from oracle_libs import load_ref      # noqa: E402

MASTER = 0x34F05C64D7AD598F
TRIALS = 64


class Tut3Out(C.Structure):
    _fields_ = [("events", C.c_uint64), ("t_end", C.c_double), ("park", C.c_double), ("riding", C.c_double),
                ("waiting", C.c_double), ("walking", C.c_double), ("rides", C.c_double)]


def main():
    ref = load_ref()
    assert ref is not None
    lib = C.CDLL(str(ROOT / "oracle/_ref/libtut3_ref.so"))
    lib.tut3_ref_trial.argtypes = [C.c_uint64, C.POINTER(Tut3Out)]
    rows = []
    for i in range(TRIALS):
        o = Tut3Out()
        assert lib.tut3_ref_trial(ref.ref_fmix64(MASTER, i), C.byref(o)) == 0
        rows.append({"events": o.events, "t_end": float(o.t_end).hex(), "means": [float(v).hex() for v in (o.park, o.riding, o.waiting, o.walking, o.rides)]})
    (ROOT / "tests/golden/park_vectors.json").write_text(json.dumps({"master": MASTER, "trials": rows}, indent=1))
    print(TRIALS, "trials,", sum(r["events"] for r in rows), "events")


if __name__ == "__main__":
    main()
