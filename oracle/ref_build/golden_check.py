#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.

Runs the reference's own golden-output programs (test/test_<name>.c, built by
`make -C oracle golden-build` against the unmodified reference sources) with the
reference's fixed seed and compares stdout with test/reference/<name>.txt using
the same normalisation rule as the reference's test/tools/test_stochastic.py:92-105
(decode, split on universal newlines, drop trailing blank lines).

usage: golden_check.py <reference root> <dir with test_<name> binaries>
"""
import subprocess
import sys
from pathlib import Path

SEED = 0x34F05C64D7AD598F      # test/tools/test_stochastic.py:58-69
NAMES = ["buffer", "condition", "data", "event", "hashheap", "objectqueue",
         "priorityqueue", "random", "resource", "resourcepool"]


def lines_of(raw: bytes):
    out = raw.decode("utf-8", errors="replace").splitlines()
    while out and not out[-1].strip():
        out.pop()
    return out


def main() -> int:
    ref_root, bindir = Path(sys.argv[1]), Path(sys.argv[2])
    bad = 0
    for name in NAMES:
        proc = subprocess.run([str(bindir / f"test_{name}"), "-s", str(SEED)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
        want = lines_of((ref_root / "test" / "reference" / f"{name}.txt").read_bytes())
        got = lines_of(proc.stdout)
        good = proc.returncode == 0 and got == want
        print(f"{'PASS' if good else 'FAIL'} {name}: {len(got)} lines, exit {proc.returncode}")
        bad += 0 if good else 1
    print("golden files matched:", len(NAMES) - bad, "of", len(NAMES))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
