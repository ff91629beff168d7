#!/usr/bin/env python3
"""Regenerate the committed golden vectors from the UNMODIFIED reference.

Runs in the build container only (needs oracle/_ref, i.e. /root/reference and
`make -C oracle ref`).  Everything written here comes out of the reference's own
code: cmb_random_* streams, seeded runs of the benchmark model through
cmb_event_execute_next(), and cmb_datasummary/cmb_wtdsummary add + merge.

Doubles are stored as C99 hex-float strings (exact).  Long streams are stored as
the first values verbatim plus XOR / wrapping-sum checksums of the 64-bit
patterns of every value, so a single flipped bit anywhere in 10^6 draws fails.

usage: python tests/golden/make_golden.py
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "tests"))
from oracle_libs import DIST_CASES, rng_draws_ex, load_ref, Result, trace_trial, rng_draws   # noqa: E402

KAT_SEED = 0x34F05C64D7AD598F          # test/tools/test_stochastic.py:58-69
SEEDS = [KAT_SEED, 12345, 7, 0, 2**64 - 1]
RNG_KINDS = {0: (0.0, 0.0), 1: (1.0, 0.0), 2: (0.0, 0.0), 3: (0.0, 0.0), 4: (1.0, 0.25),
             5: (2.0, 0.625), 6: (-3.0, 5.5), 7: (1.0, 6.0), 8: (0.3, 0.0)}
MODELS = {0: dict(arr=1 / 0.9, srv=1.0, servers=1), 1: dict(arr=1.25, srv=1.0, servers=1),
          2: dict(arr=1 / 6.4, srv=1.0, servers=8),
          3: dict(arr=1.0, srv=1.0, servers=10),     # model 3: num_objects = duration, servers = queue capacity
          4: dict(arr=1.0, srv=1.0, servers=20),     # model 4: num_objects = duration, servers = pool capacity
          5: dict(arr=1.0, srv=1.0, servers=10),     # model 5: num_objects = duration, servers = buffer capacity
          6: dict(arr=1.0, srv=1.0, servers=8),      # model 6: num_objects = duration, servers = queue capacity
          7: dict(arr=1.0, srv=1.0, servers=200),    # model 7: num_objects = duration, servers = worker processes
          8: dict(arr=1.0, srv=0.6, servers=1),      # model 8: num_objects = duration
          9: dict(arr=1 / 0.9, srv=1.0, servers=1),  # model 9: M/M/1 with the queue history on (counters = wtdsummary bits)
          10: dict(arr=2.0, srv=8.0, servers=10),    # model 10: the harbor of test/test_condition.c (num_objects = hours)
          11: dict(arr=1.0, srv=1.0, servers=10),    # model 11: test/test_objectqueue.c with the queue history on
          12: dict(arr=1.0, srv=1.0, servers=10),    # model 12: test/test_buffer.c as it stands
          13: dict(arr=1.0, srv=1.0, servers=10),    # model 13: test/test_priorityqueue.c
          14: dict(arr=1.0, srv=1.0, servers=1)}     # model 14: test/test_resource.c as it stands
# the reference's own golden runs (test/reference/*.txt) that a model reproduces: model -> (file, size)
GOLDEN_RUNS = {10: ("condition.txt", 24 * 7 * 52 * 100), 11: ("objectqueue.txt", 1_000_000),
               12: ("buffer.txt", 10_000), 13: ("priorityqueue.txt", 1_000_000), 14: ("resource.txt", 25)}
FULL_SIZE_KAT = (0, 1, 2, 9)             # 10^6 objects with the KAT seed


def hexes(a):
    return [float.hex(float(x)) for x in a]


def checksum(a):
    u = np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
    return {"xor": int(np.bitwise_xor.reduce(u)), "sum": int(np.add.reduce(u, dtype=np.uint64))}


def summary_stats(ref, x):
    """Skewness, excess kurtosis and the printed line of the reference's own cmb_datasummary for prefixes of x
    (the counts 0..4 exercise every column rule of cmb_datasummary_print, src/cmb_datasummary.c:168-212)."""
    import tempfile

    class DS(C.Structure):          # include/cmb_datasummary.h:42-51
        _fields_ = [("cookie", C.c_uint64), ("count", C.c_uint64), ("min", C.c_double), ("max", C.c_double),
                    ("m1", C.c_double), ("m2", C.c_double), ("m3", C.c_double), ("m4", C.c_double)]
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    for name in ("cmb_datasummary_skewness", "cmb_datasummary_kurtosis"):
        getattr(ref, name).restype = C.c_double
        getattr(ref, name).argtypes = [C.POINTER(DS)]
    ref.cmb_datasummary_add.argtypes = [C.POINTER(DS), C.c_double]
    ref.cmb_datasummary_print.argtypes = [C.POINTER(DS), C.c_void_p, C.c_bool]
    stats = {}
    for n in (0, 1, 2, 3, 4, 10, 1000):
        ds = DS()
        ref.cmb_datasummary_initialize(C.byref(ds))
        for v in x[:n]:
            ref.cmb_datasummary_add(C.byref(ds), float(v))
        lines = []
        for lead in (True, False):
            with tempfile.NamedTemporaryFile() as tmp:
                fp = libc.fopen(tmp.name.encode(), b"w")
                ref.cmb_datasummary_print(C.byref(ds), fp, lead)
                libc.fclose(fp)
                lines.append(open(tmp.name).read())
        stats[str(n)] = {"skewness": float.hex(ref.cmb_datasummary_skewness(C.byref(ds))),
                         "kurtosis": float.hex(ref.cmb_datasummary_kurtosis(C.byref(ds))),
                         "line": lines[0], "line_plain": lines[1]}
    return stats


def awacs_vectors():
    """tests/golden/awacs_vectors.json from the unmodified tutorial source (oracle/_ref/libawacs_ref.so)."""
    import hashlib
    from oracle_libs import AWACS_TERRAIN_SEED, awacs_terrain, awacs_trial, load_awacs_ref
    ref = load_awacs_ref()
    if ref is None:
        sys.exit("oracle/_ref/libawacs_ref.so is not built; run `make -C oracle ref` first")
    width, height, duration = 12.0, 10.0, 0.05
    m, cols, rows, geom = awacs_terrain(ref, "ref", AWACS_TERRAIN_SEED, width, height)
    out = {"source": "tutorial/tut_5_1.c, unmodified, behind oracle/ref_build/awacs_stubs/hdf5.h",
           "terrain": {"seed": AWACS_TERRAIN_SEED, "width_nm": width, "height_nm": height, "cols": cols, "rows": rows,
                       "geom": [float(v).hex() for v in geom], "map_sha256": hashlib.sha256(m.tobytes()).hexdigest()},
           "duration_h": duration, "trials": [], "platform": {}}
    for i in range(4):
        seed = int(ref_fmix(KAT_SEED, i))
        o, keys, times, per = awacs_trial(ref, "ref", seed, duration, trace_cap=4000)
        trace = hashlib.sha256(np.array(keys, dtype=np.uint64).tobytes() + np.array(times, dtype=np.float64).tobytes())
        out["trials"].append({"seed": seed, "events": o.events, "t_end": o.t_end.hex(), "num_found": o.num_found,
                              "tds_count": list(o.tds_count), "mode_count": list(o.mode_count),
                              "sum_x": o.sum_x.hex(), "sum_y": o.sum_y.hex(), "trace_sha256": trace.hexdigest(),
                              "tds_sha256": hashlib.sha256(np.array(per["tds"], dtype=np.int32).tobytes()).hexdigest()})
    six, r = (C.c_float * 6)(), C.c_float()
    for t in (0.0, 1.0, 150.0, 700.5, 900.0, 1400.25, 2000.0, 86399.0):
        ref.awacs_ref_platform_state(C.c_double(t), six, C.byref(r))
        out["platform"][repr(t)] = [float(v).hex() for v in six] + [float(r.value).hex()]
    path = ROOT / "tests/golden/awacs_vectors.json"
    path.write_text(json.dumps(out, indent=0) + "\n")
    print("wrote", path, path.stat().st_size, "bytes")


def ref_fmix(seed, nonce):
    lib = load_ref()
    return lib.ref_fmix64(seed, nonce)


def main():
    if "--only-awacs" in sys.argv:
        awacs_vectors()
        return
    ref = load_ref()
    if "--only-summary-stats" in sys.argv:      # refresh one block; everything else stays as committed
        path = ROOT / "tests/golden/reference_vectors.json"
        out = json.loads(path.read_text())
        x = np.array([float.fromhex(v) for v in out["summary"]["x"]])
        out["summary_stats"] = summary_stats(ref, x)
        path.write_text(json.dumps(out, indent=0) + "\n")
        print("updated summary_stats in", path)
        return
    if ref is None:
        sys.exit("oracle/_ref is not built; run `make -C oracle ref` first")
    out = {"source": "ambonvik/cimba via oracle/_ref (unmodified reference build)"}

    rng = {}
    for seed in SEEDS[:3]:
        per = {}
        for kind, (p0, p1) in RNG_KINDS.items():
            n = 1_000_000 if kind in (0, 1, 2, 4, 5) else 100_000
            v = rng_draws(ref, "ref", seed, kind, p0, p1, n)
            if kind == 0:
                first = [int(x) for x in v[:16].view(np.uint64)]
            else:
                first = hexes(v[:16])
            per[str(kind)] = {"p0": p0, "p1": p1, "n": n, "first": first, **checksum(v)}
        rng[str(seed)] = per
    out["rng"] = rng

    # the rest of cmb_random: 65 536 draws per case (a multiple of 64, see rng_draws_ex), KAT seed
    dist = []
    for kind, par in DIST_CASES:
        v = np.array(rng_draws_ex(ref, "ref", KAT_SEED, kind, par, 65_536))
        dist.append({"kind": kind, "params": par, "n": 65_536, "first": hexes(v[:8]), **checksum(v)})
    out["distributions"] = dist
    out["fmix64"] = {str(s): [int(ref.ref_fmix64(s, k)) for k in range(4)] for s in SEEDS}

    trials = []
    for model, par in MODELS.items():
        for seed in SEEDS:
            # model 3: the size is a duration; 0 would stop workers before they start (they then
            # run forever in the reference), so it starts at 1
            traced = 10 if model == 7 else 1000
            for nobj in ((0, 1, 2, 3, 10, 1000, 100_000) if model in FULL_SIZE_KAT else ((1, 2, 3, 10, 50) if model == 7 else (1, 2, 3, 10, 100, 1000))):
                r, keys, times = trace_trial(ref, "ref", model, par["servers"], seed, nobj,
                                             par["arr"], par["srv"], 512 if nobj == traced else 0)
                rec = {"model": model, "servers": par["servers"], "seed": seed, "num_objects": nobj,
                       "arr_mean": float.hex(par["arr"]), "srv_mean": float.hex(par["srv"]),
                       "events": r.events, "objects": r.objects, "t_end": float.hex(r.t_end),
                       "sum_wait": float.hex(r.sum_wait), "max_fel": r.max_fel, "max_queue": r.max_queue,
                       "counters": r.counters()}
                if nobj == traced:
                    rec["trace_key"] = [int(k) for k in keys]
                    rec["trace_time"] = hexes(times)
                trials.append(rec)
        big = GOLDEN_RUNS[model][1] if model in GOLDEN_RUNS else (1_000_000 if model in FULL_SIZE_KAT else None)
        if big is None:
            continue
        # one full-size record with the KAT seed: the reference's golden run / SURVEY.md section 8c's known answer
        r, _, _ = trace_trial(ref, "ref", model, par["servers"], KAT_SEED, big, par["arr"], par["srv"], 0)
        trials.append({"model": model, "servers": par["servers"], "seed": KAT_SEED, "num_objects": big,
                       "arr_mean": float.hex(par["arr"]), "srv_mean": float.hex(par["srv"]),
                       "events": r.events, "objects": r.objects, "t_end": float.hex(r.t_end),
                       "sum_wait": float.hex(r.sum_wait), "max_fel": r.max_fel, "max_queue": r.max_queue,
                       "counters": r.counters()})
    out["trials"] = trials

    # experiment-level: first 64 trials of the fmix64-seeded M/M/1 experiment, 10 000 objects
    n = 64
    res = (Result * n)()
    ref.ref_run_trials(0, 1, KAT_SEED, 0, n, 10_000, 1 / 0.9, 1.0, 0, res)
    out["experiment_mm1"] = {"master_seed": KAT_SEED, "num_objects": 10_000, "trials": [
        {"events": r.events, "objects": r.objects, "t_end": float.hex(r.t_end), "sum_wait": float.hex(r.sum_wait)}
        for r in res]}

    # summaries
    g = np.random.default_rng(1234)
    x = g.gamma(2.0, 3.0, size=1000)
    w = g.uniform(0.1, 2.0, size=1000)
    o7, o8 = (C.c_double * 8)(), (C.c_double * 8)()
    xs = x.ctypes.data_as(C.POINTER(C.c_double))
    wsp = w.ctypes.data_as(C.POINTER(C.c_double))
    summ = {"x": hexes(x), "w": hexes(w)}
    ref.ref_datasummary_of(xs, 1000, o7); summ["data_all"] = hexes(o7[:7])
    ref.ref_wtdsummary_of(xs, wsp, 1000, o8); summ["wtd_all"] = hexes(o8[:8])
    for na in (1, 333, 500, 999):
        ref.ref_datasummary_split_merge(xs, na, 1000, o7); summ[f"data_merge_{na}"] = hexes(o7[:7])
        ref.ref_wtdsummary_split_merge(xs, wsp, na, 1000, o8); summ[f"wtd_merge_{na}"] = hexes(o8[:8])
    out["summary"] = summ
    out["summary_stats"] = summary_stats(ref, x)

    path = ROOT / "tests/golden/reference_vectors.json"
    path.write_text(json.dumps(out, indent=0) + "\n")
    print("wrote", path, path.stat().st_size, "bytes")
    awacs_vectors()


if __name__ == "__main__":
    main()
