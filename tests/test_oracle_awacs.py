"""CPU tests of the AWACS oracle (tutorial/tut_5_1.c, BASELINE config 5): the plain-C restatement
(oracle/port/awacs_port.c) against the UNMODIFIED tutorial source compiled behind a stub hdf5.h
(oracle/_ref/libawacs_ref.so) wherever that build is present, and against the vectors that build produced
(tests/golden/awacs_vectors.json, tests/golden/make_golden.py --only-awacs) everywhere."""
import ctypes as C
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from oracle_libs import (AWACS_TERRAIN_SEED, awacs_terrain, awacs_trial, load_awacs_ref, load_port)

GOLD = json.loads((Path(__file__).parent / "golden/awacs_vectors.json").read_text())


@pytest.fixture(scope="module")
def port():
    return load_port()


@pytest.fixture(scope="module")
def port_terrain(port):
    g = GOLD["terrain"]
    return awacs_terrain(port, "port", AWACS_TERRAIN_SEED, g["width_nm"], g["height_nm"])


def test_port_terrain_matches_the_reference_vectors(port_terrain):
    m, cols, rows, geom = port_terrain
    g = GOLD["terrain"]
    assert (cols, rows) == (g["cols"], g["rows"])
    assert [float(v).hex() for v in geom] == g["geom"]
    assert hashlib.sha256(m.tobytes()).hexdigest() == g["map_sha256"]


def test_threaded_terrain_generation_gives_the_same_map(port, port_terrain):
    g = GOLD["terrain"]
    m, cols, rows, geom = awacs_terrain(port, "port", AWACS_TERRAIN_SEED, g["width_nm"], g["height_nm"], threads=5)
    assert (cols, rows) == port_terrain[1:3] and np.array_equal(geom, port_terrain[3])
    assert np.array_equal(m.view(np.uint32), port_terrain[0].view(np.uint32))


@pytest.mark.parametrize("case", GOLD["trials"], ids=lambda c: f"seed{c['seed']}")
def test_port_trial_matches_the_reference_vectors(port, port_terrain, case):
    out, keys, times, per = awacs_trial(port, "port", case["seed"], GOLD["duration_h"], port_terrain, trace_cap=4000)
    assert out.events == case["events"] and out.t_end.hex() == case["t_end"] and out.num_found == case["num_found"]
    assert list(out.tds_count) == case["tds_count"] and list(out.mode_count) == case["mode_count"]
    assert out.sum_x.hex() == case["sum_x"] and out.sum_y.hex() == case["sum_y"]
    trace = hashlib.sha256(np.array(keys, dtype=np.uint64).tobytes() + np.array(times, dtype=np.float64).tobytes())
    assert trace.hexdigest() == case["trace_sha256"]
    assert hashlib.sha256(np.array(per["tds"], dtype=np.int32).tobytes()).hexdigest() == case["tds_sha256"]


def test_port_platform_state_matches_the_reference_vectors(port):
    six, r = (C.c_float * 6)(), C.c_float()
    for t, want in GOLD["platform"].items():
        port.port_awacs_platform_state(C.c_double(float(t)), six, C.byref(r))
        assert [float(v).hex() for v in six] + [float(r.value).hex()] == want


def test_port_matches_the_live_reference_build(port):
    ref = load_awacs_ref()
    if ref is None:
        pytest.skip("oracle/_ref/libawacs_ref.so not built (needs /root/reference)")
    rt = awacs_terrain(ref, "ref", 77, 8.0, 6.0)
    pt = awacs_terrain(port, "port", 77, 8.0, 6.0)
    assert rt[1:3] == pt[1:3] and np.array_equal(rt[3], pt[3])
    assert np.array_equal(rt[0].view(np.uint32), pt[0].view(np.uint32))
    for seed in (5, 6):
        ro, rk, rtm, rper = awacs_trial(ref, "ref", seed, 0.03, trace_cap=3000)
        po, pk, ptm, pper = awacs_trial(port, "port", seed, 0.03, pt, trace_cap=3000)
        assert ro.key() == po.key() and rk == pk and rtm == ptm
        assert rper["tds"] == pper["tds"] and rper["detected"] == pper["detected"] and rper["mode"] == pper["mode"]
        assert np.array_equal(rper["x"].view(np.uint32), pper["x"].view(np.uint32))


def test_reference_executive_runs_awacs_trials_in_parallel(port):
    """cimba_run_experiment over the tutorial's run_trial (all host cores) gives what one-at-a-time runs give."""
    from oracle_libs import awacs_ref_experiment, load_ref
    ref = load_awacs_ref()
    if ref is None:
        pytest.skip("oracle/_ref/libawacs_ref.so not built (needs /root/reference)")
    pt = awacs_terrain(port, "port", 77, 8.0, 6.0)
    awacs_terrain(ref, "ref", 77, 8.0, 6.0)
    master = 0x34F05C64D7AD598F
    outs = awacs_ref_experiment(ref, master, 3, 6, 0.02)
    fmix = load_ref().ref_fmix64
    for i, o in enumerate(outs):
        po, _, _, _ = awacs_trial(port, "port", fmix(master, 3 + i), 0.02, pt)
        assert o.key() == po.key()
