#!/usr/bin/env python
"""List the local-memory instructions (LDL/STL: spills, stack arrays) of one kernel with their source lines.

    python tools/sass_local.py <cubin> <mangled-kernel-regex>
"""
import re
import subprocess
import sys

cubin, kre = sys.argv[1], sys.argv[2]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
on, line, inl = False, 0, ""
n = 0
for l in txt.splitlines():
    m = re.match(r"\.text\.(\S+):", l)
    if m:
        on = re.search(kre, m.group(1)) is not None
        if on:
            print("==", m.group(1))
        continue
    m = re.search(r'//## File ".*", line (\d+)(.*)', l)
    if m:
        line, inl = int(m.group(1)), m.group(2)
        continue
    if on and re.search(r"\b(LDL|STL)", l):
        n += 1
        print(f"{line:5d} {l.strip()[:90]} {inl[:60]}")
print(n, "local-memory instructions")
