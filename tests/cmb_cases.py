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


def check_trial(want, events, objects, t_end, sum_wait, counters, keys, times, what=""):
    assert int(events) == want["events"], (what, "events", int(events), want["events"])
    assert int(objects) == want["objects"], (what, "objects")
    assert float(t_end).hex() == want["t_end"], (what, "t_end", float(t_end), float.fromhex(want["t_end"]))
    assert float(sum_wait).hex() == want["sum_wait"], (what, "sum_wait")
    if counters is not None and any(want["counters"]):
        assert [int(v) for v in counters[:4]] == want["counters"], (what, "counters")
    if keys is not None:
        assert trace_digest(keys, times, events) == want["trace_sha256"], (what, "pop trace")
