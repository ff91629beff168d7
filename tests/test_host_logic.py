"""CPU tests of the host-side logic: cmb_datasummary mirror (vs the reference's
golden numbers and the oracle), and the cross-rank merge over gloo, world_size 2."""
import ctypes as C
import math
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_datasummary_add_matches_reference_golden(cb, golden):
    s = golden["summary"]
    x = [float.fromhex(v) for v in s["x"]]
    ds = cb.DataSummary.of(x)
    assert [float.hex(v) for v in ds.to_list()[:7]] == s["data_all"]
    assert ds.count() == 1000 and ds.mean() == float.fromhex(s["data_all"][3])
    assert ds.variance() == float.fromhex(s["data_all"][4]) / 999.0


@pytest.mark.parametrize("na", [1, 333, 500, 999])
def test_datasummary_merge_matches_reference_golden(cb, golden, na):
    s = golden["summary"]
    x = [float.fromhex(v) for v in s["x"]]
    m = cb.DataSummary.merge(cb.DataSummary.of(x[:na]), cb.DataSummary.of(x[na:]))
    assert [float.hex(v) for v in m.to_list()[:7]] == s[f"data_merge_{na}"]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 10, 1000])
def test_skewness_kurtosis_and_print_match_reference_golden(cb, golden, n):
    """cmb_datasummary_skewness / _kurtosis / _print (src/cmb_datasummary.c:168-249) of the reference itself,
    recorded by tests/golden/make_golden.py: same bits, same text, column rules included (n = 0..4)."""
    x = [float.fromhex(v) for v in golden["summary"]["x"]]
    want = golden["summary_stats"][str(n)]
    ds = cb.DataSummary.of(x[:n])
    assert float.hex(ds.skewness()) == want["skewness"]
    assert float.hex(ds.kurtosis()) == want["kurtosis"]
    assert ds.line(True) == want["line"] and ds.line(False) == want["line_plain"]
    ws = cb.WtdSummary()
    for v in x[:n]:
        ws.add(v, 1.0)
    assert ws.count() == n and ws.line(True).startswith("N ")
    if n > 3:
        assert math.isfinite(ws.skewness()) and math.isfinite(ws.kurtosis()) and ws.stddev() > 0.0


def test_merge_with_empty_side_and_roundtrip(cb):
    a = cb.DataSummary.of([1.0, 2.0, 4.0])
    e = cb.DataSummary()
    for m in (cb.DataSummary.merge(a, e), cb.DataSummary.merge(e, a)):
        assert m.to_list() == a.to_list()
    assert cb.DataSummary.from_list(a.to_list()).to_list() == a.to_list()
    assert math.isnan(cb.DataSummary.of([3.0]).half_width_95())
    assert e.count() == 0 and e.variance() == 0.0


def test_tree_merge_is_within_tolerance_of_serial_fold(cb):
    """A merged tree differs from the reference's serial add loop by rounding only
    (<< 1e-9 relative, the BASELINE.json tolerance)."""
    g = np.random.default_rng(7)
    x = g.gamma(3.0, 3.3, size=4096)
    serial = cb.DataSummary.of(x)
    parts = [cb.DataSummary.of(x[i::8]) for i in range(8)]
    acc = parts[0]
    for p in parts[1:]:
        acc = cb.DataSummary.merge(acc, p)
    assert acc.count() == serial.count() and acc.min() == serial.min() and acc.max() == serial.max()
    for a, b in zip(acc.to_list()[3:7], serial.to_list()[3:7]):
        assert abs(a - b) <= 1e-9 * abs(b)


def test_merge_across_ranks_without_process_group(cb):
    import torch
    ds = cb.DataSummary.of([1.0, 5.0, 9.0])
    out = cb.merge_across_ranks(torch.tensor(ds.to_list(), dtype=torch.float64))
    assert out.to_list() == ds.to_list()


def _wtd_of(cb, x, w):
    out = cb.WtdSummary()
    for xi, wi in zip(x, w):
        out.add(xi, wi)
    return out


def test_wtdsummary_add_and_merge_match_reference_golden(cb, golden):
    """cmb_wtdsummary_add / _merge of the C-ABI library vs numbers produced by the reference itself."""
    s = golden["summary"]
    x = [float.fromhex(v) for v in s["x"]]
    w = [float.fromhex(v) for v in s["w"]]
    assert [float.hex(v) for v in _wtd_of(cb, x, w).fields()] == s["wtd_all"]
    for na in (1, 333, 500, 999):
        m = cb.WtdSummary.merge(_wtd_of(cb, x[:na], w[:na]), _wtd_of(cb, x[na:], w[na:]))
        assert [float.hex(v) for v in m.fields()] == s[f"wtd_merge_{na}"]
    a = _wtd_of(cb, x[:10], w[:10])
    assert cb.WtdSummary.from_row(a.to_row()).fields() == a.fields()
    assert a.variance() == a.fields()[4] / 9.0        # include/cmb_wtdsummary.h:192-197: the base-class variance
    e = cb.WtdSummary()
    assert cb.WtdSummary.merge(a, e).fields() == a.fields() == cb.WtdSummary.merge(e, a).fields()
    z = cb.WtdSummary()
    assert z.add(3.0, 0.0) == 0 and z.count() == 0    # zero weight: ignored (src/cmb_wtdsummary.c:92-94)


def test_alias_tables_match_the_oracle(cb, port):
    """cimba_b200_alias_create (host code of the C-ABI library) vs the oracle's restatement of
    cmb_random_alias_create (src/cmb_random.c:688-752), which test_oracle_port pins to the reference."""
    for probs in ([0.05, 0.25, 0.4, 0.1, 0.2], [1.0], [0.5, 0.5], [0.999, 0.0005, 0.0005], [1 / 7] * 7,
                  [0.01 * k for k in range(1, 14)] + [0.09]):
        n = len(probs)
        uprob, alias = cb.alias_create(probs)
        pa = (C.c_double * n)(*probs)
        pu = (C.c_uint64 * n)()
        pal = (C.c_uint * n)()
        port.port_alias_create.restype = None
        port.port_alias_create.argtypes = [C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint)]
        port.port_alias_create(n, pa, pu, pal)
        assert uprob == list(pu) and alias == list(pal), probs
    with pytest.raises(cb.CimbaError):
        cb.alias_create([0.5, 0.2])                    # does not sum to one (src/cmb_random.c:634-642)


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    import cimba_b200 as cb
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
    rank = dist.get_rank()
    x = np.random.default_rng(11).gamma(2.0, 5.0, size=2000)
    shard = x[rank * 1000:(rank + 1) * 1000]          # contiguous blocks of the trial array, section 8e
    local = torch.tensor(cb.DataSummary.of(shard).to_list(), dtype=torch.float64)
    merged = cb.merge_across_ranks(local)
    w = np.random.default_rng(12).uniform(0.1, 2.0, size=2000)[rank * 1000:(rank + 1) * 1000]
    mine = cb.WtdSummary()
    for xi, wi in zip(shard, w):
        mine.add(xi, wi)
    wmerged = cb.merge_weighted_across_ranks(torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in mine.to_row()],
                                                          dtype=torch.int64))
    print(json.dumps({{"rank": rank, "summary": [float.hex(v) for v in merged.to_list()],
                      "wsummary": [float.hex(v) for v in wmerged.fields()]}}), flush=True)
    dist.destroy_process_group()
""")


def test_two_rank_gloo_merge_equals_single_process_merge(cb, tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT), port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=180)
        assert p.returncode == 0, e
        outs.append(o.strip().splitlines()[-1])
    import json
    got = [json.loads(o)["summary"] for o in outs]
    assert got[0] == got[1]                          # every rank ends with the same merged summary
    x = np.random.default_rng(11).gamma(2.0, 5.0, size=2000)
    want = cb.DataSummary.merge(cb.DataSummary.of(x[:1000]), cb.DataSummary.of(x[1000:]))
    assert got[0] == [float.hex(v) for v in want.to_list()]
    serial = cb.DataSummary.of(x)
    assert abs(want.mean() - serial.mean()) <= 1e-12 * abs(serial.mean())
    # the weighted half of the same exchange: cmb_wtdsummary rows all-gathered and merged in rank order
    wgot = [json.loads(o)["wsummary"] for o in outs]
    assert wgot[0] == wgot[1]
    w = np.random.default_rng(12).uniform(0.1, 2.0, size=2000)
    wwant = cb.WtdSummary.merge(_wtd_of(cb, x[:1000], w[:1000]), _wtd_of(cb, x[1000:], w[1000:]))
    assert wgot[0] == [float.hex(v) for v in wwant.fields()]


def test_bench_reference_arm_prints_one_contract_line(tmp_path):
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm): exactly one JSON line on
    stdout with the contract's keys, whatever the libraries print elsewhere."""
    import json
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--ref-trials", "8", "--objects", "2000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "events/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]
