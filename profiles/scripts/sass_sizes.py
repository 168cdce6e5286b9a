#!/usr/bin/env python
"""SASS instruction count (and code bytes) per kernel of the built library: python profiles/scripts/sass_sizes.py [lib]"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "conclave_sim_b200/libconclave_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, cnt = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        cnt[name] = 0
    elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        cnt[name] += 1
names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
for (k, v), n in sorted(zip(cnt.items(), names), key=lambda t: -t[0][1]):
    print(f"{v:7d} instr {v * 16 / 1024:7.1f} KB  {n}")
