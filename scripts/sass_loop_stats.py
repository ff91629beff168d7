#!/usr/bin/env python3
"""Static view of a kernel's main loop in SASS: instruction count and opcode mix between the
first DSETP (top of the event step) and the last backward branch.
usage: sass_loop_stats.py <lib.so> <kernel-name-substring>"""
import collections
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
body = [b for b in blocks if pat in b.split("\n", 1)[0]][0]
ins = []
for l in body.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
back = [(a, t) for a, t in ins if re.search(r"\bBRA\b", t) and int(re.search(r"0x([0-9a-f]+)", t).group(1), 16) < a]
# the main loop = the backward branch with the largest span
a_end, t_end = max(back, key=lambda x: x[0] - int(re.search(r"0x([0-9a-f]+)", x[1]).group(1), 16))
a_start = int(re.search(r"0x([0-9a-f]+)", t_end).group(1), 16)
loop = [(a, t) for a, t in ins if a_start <= a <= a_end]
mix = collections.Counter()
for a, t in loop:
    op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
    mix[op] += 1
print(f"loop 0x{a_start:x}..0x{a_end:x}: {len(loop)} instructions (all paths, incl. rare blocks)")
print(", ".join(f"{k} {v}" for k, v in mix.most_common()))
