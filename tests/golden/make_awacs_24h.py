"""Golden vectors for FULL 24-hour AWACS trials (tutorial/tut_5_1.c:1211-1213, BASELINE config 5), produced by the
UNMODIFIED tutorial source (oracle/_ref/libawacs_ref.so = tut_5_1.c behind a stub hdf5.h, glibc libm).

    python tests/golden/make_awacs_24h.py [--trials 8] [--width 100 --height 100] [--hours 24] [--procs 8]

One 24-hour trial is 8.64e7 target sweeps and ~1.1e5 events: about 25 minutes on one host core, so the trials run
as independent processes (each builds the same terrain from AWACS_TERRAIN_SEED, then runs its trial; the reference
keeps the terrain in process-global state).  Output: tests/golden/awacs_24h.npz with, per trial i (seed
cmb_random_fmix64(MASTER, first + i)): events, t_end, num_found, tds_count[6], mode_count[4], sum_x, sum_y and the
1000 final per-target x, y (float32 bit patterns), mode, detect state and found flag.  The GPU test
tests/test_gpu_awacs.py::test_awacs_full_24_hour_trials_match_the_reference compares the device against these.
"""
import argparse
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
MASTER = 0x34F05C64D7AD598F


def one(args):
    i, first, width, height, hours = args
    from oracle_libs import AWACS_TERRAIN_SEED, awacs_terrain, awacs_trial, load_awacs_ref, load_ref
    ref = load_awacs_ref()
    assert ref is not None, "oracle/_ref/libawacs_ref.so is not built (make -C oracle ref)"
    seed = load_ref().ref_fmix64(MASTER, first + i)
    m, cols, rows, geom = awacs_terrain(ref, "ref", AWACS_TERRAIN_SEED, width, height)
    t0 = time.time()
    out, _, _, per = awacs_trial(ref, "ref", seed, hours)
    print(f"trial {first + i}: {out.events} events, {out.num_found} found, {time.time() - t0:.0f} s", flush=True)
    return dict(i=i, seed=seed, cols=cols, rows=rows, events=out.events, t_end=out.t_end, num_found=out.num_found,
                tds_count=list(out.tds_count), mode_count=list(out.mode_count), sum_x=out.sum_x, sum_y=out.sum_y,
                x=per["x"], y=per["y"], mode=per["mode"], tds=per["tds"], det=per["detected"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=8)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--width", type=float, default=100.0)
    ap.add_argument("--height", type=float, default=100.0)
    ap.add_argument("--hours", type=float, default=24.0)
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--out", default=str(ROOT / "tests/golden/awacs_24h.npz"))
    a = ap.parse_args()
    with mp.get_context("spawn").Pool(a.procs) as pool:
        rows = pool.map(one, [(i, a.first, a.width, a.height, a.hours) for i in range(a.trials)], chunksize=1)
    rows.sort(key=lambda r: r["i"])
    np.savez_compressed(
        a.out,
        master=np.uint64(MASTER), first=np.uint64(a.first), width_nm=np.float64(a.width), height_nm=np.float64(a.height),
        hours=np.float64(a.hours), grid=np.array([rows[0]["cols"], rows[0]["rows"]], dtype=np.uint32),
        seed=np.array([r["seed"] for r in rows], dtype=np.uint64),
        events=np.array([r["events"] for r in rows], dtype=np.uint64),
        t_end=np.array([r["t_end"] for r in rows], dtype=np.float64),
        num_found=np.array([r["num_found"] for r in rows], dtype=np.uint32),
        tds_count=np.array([r["tds_count"] for r in rows], dtype=np.uint32),
        mode_count=np.array([r["mode_count"] for r in rows], dtype=np.uint32),
        sum_x=np.array([r["sum_x"] for r in rows], dtype=np.float64),
        sum_y=np.array([r["sum_y"] for r in rows], dtype=np.float64),
        x_bits=np.stack([r["x"].view(np.uint32) for r in rows]),
        y_bits=np.stack([r["y"].view(np.uint32) for r in rows]),
        mode=np.array([r["mode"] for r in rows], dtype=np.uint8),
        tds=np.array([r["tds"] for r in rows], dtype=np.uint8),
        det=np.array([r["det"] for r in rows], dtype=np.uint8))
    print("wrote", a.out)


if __name__ == "__main__":
    main()
