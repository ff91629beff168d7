"""Shared by tests/test_cmb_engine.py (CPU) and tests/test_gpu_cmb_engine.py (GPU): the golden cases of the general
engine's models (tests/golden/cmb_engine_vectors.json, produced by the unmodified reference) and how a result row is
compared with them."""
import hashlib
import json
from pathlib import Path

import numpy as np

GOLD = json.loads((Path(__file__).parent / "golden/cmb_engine_vectors.json").read_text())
MASTER = GOLD["master"]
TRACE = GOLD["trace"]


def case_id(c):
    return f"model{c['model']}-s{c['servers']}-n{c['num_objects']}-a{float.fromhex(c['arr_mean']):.3f}"


def trace_digest(keys, times, pops):
    n = min(int(pops), TRACE)
    return hashlib.sha256(np.asarray(keys[:n], dtype=np.uint64).tobytes() + np.asarray(times[:n], dtype=np.float64).tobytes()).hexdigest()


def check_trial(want, events, objects, t_end, sum_wait, counters, keys, times, what="", max_queue=None):
    assert int(events) == want["events"], (what, "events", int(events), want["events"])
    assert int(objects) == want["objects"], (what, "objects")
    assert float(t_end).hex() == want["t_end"], (what, "t_end", float(t_end), float.fromhex(want["t_end"]))
    assert float(sum_wait).hex() == want["sum_wait"], (what, "sum_wait")
    if counters is not None and any(want["counters"]):
        assert [int(v) for v in counters[:4]] == want["counters"], (what, "counters")
        if want.get("counters8") and int(want.get("all8", 0)):
            assert [int(v) & (2**64 - 1) for v in counters[:8]] == want["counters8"], (what, "all eight counters")
    if max_queue is not None:                           # the callers pass it for the models whose max_queue is a history's sample count
        assert int(max_queue) == want["max_queue"], (what, "history samples", int(max_queue), want["max_queue"])
    if keys is not None:
        assert trace_digest(keys, times, events) == want["trace_sha256"], (what, "pop trace")


# test/reference/resourcepool.txt:13 - the summary line of the reference's own pool test (seed 0x34f05c64d7ad598f, 20 units, 100 time units)
RESOURCEPOOL_GOLDEN_LINE = "N      120  Mean    19.77  StdDev    1.147  Variance    1.316  Skewness   -6.626  Kurtosis    46.75"


def inverse_fmix64(y):
    """The master seed whose trial 0 is seeded with y: cmb_random_fmix64(master, 0) == y (both steps of the mixer are bijections)."""
    M = 2**64 - 1
    def unshift(v):
        return v ^ (v >> 33)
    v = unshift(y)
    v = (v * pow(0xc4ceb9fe1a85ec53, -1, 2**64)) & M
    v = unshift(v)
    v = (v * pow(0xff51afd7ed558ccd, -1, 2**64)) & M
    return unshift(v)


def wtdsummary_line(lib, counters):
    """The line cmb_wtdsummary_print writes for the eight exported words (through the C-ABI's own printer)."""
    import ctypes as C
    import os
    import struct
    import tempfile
    from cimba_b200 import _lib
    f = lambda u: struct.unpack("<d", struct.pack("<Q", int(u) & (2**64 - 1)))[0]
    ws = _lib.WtdSummaryStruct()
    lib.cimba_b200_wtdsummary_initialize(C.byref(ws))
    ws.base.count = int(counters[0])
    ws.base.min, ws.base.max, ws.base.m1, ws.base.m2, ws.base.m3, ws.base.m4 = (f(c) for c in counters[1:7])
    ws.wsum = f(counters[7])
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "line.txt")
        fp = libc.fopen(path.encode(), b"w")
        lib.cimba_b200_wtdsummary_print(C.byref(ws), C.c_void_p(fp), 1)
        libc.fclose(C.c_void_p(fp))
        return open(path).read().strip()
