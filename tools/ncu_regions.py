#!/usr/bin/env python
"""Sum an ncu source-page capture over source-line REGIONS (functions) of hider_kernels.cu.

    NCU_CUBIN=... NCU_SRC=... python tools/ncu_regions.py <report.ncu-rep> <kernel-regex>

Regions are the top-level function definitions of the source (a line belongs to the last definition that starts at or
before it); prints share of warp-stall samples, of executed warp instructions and average active lanes.
"""
import csv
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines  # noqa: E402


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[hi[0]]
    body = rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))]
    si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    dis = ncu_lines.disasm_lines(kre)
    fn = [k for k in dis if len(dis[k]) == len(body)]
    lines = dis[fn[0]]
    srcp = os.environ.get("NCU_SRC") or os.path.join(ncu_lines.ROOT, "aqsis_b200", "csrc", "hider_kernels.cu")
    src = open(srcp).read().splitlines()
    starts = []
    for n, l in enumerate(src, 1):
        m = re.match(r"(?:template<[^>]*>\s*)?(?:static\s+)?__(?:device|global|host)__.*?\b(\w+)\s*\(", l)
        if m and not l.startswith("\t"):
            starts.append((n, m.group(1)))
    def region(ln):
        name = "?"
        for n, nm in starts:
            if n <= ln:
                name = nm
            else:
                break
        return name
    agg = {}
    ts = tinst = 0
    for (ln, _), r in zip(lines, body):
        a = agg.setdefault(region(ln), [0, 0, 0])
        a[0] += int(r[si]); a[1] += int(r[ii]); a[2] += int(r[ti])
        ts += int(r[si]); tinst += int(r[ii])
    print(f"{'region':28s} {'samples%':>8} {'inst%':>7} {'Ginst':>7} {'lanes':>5}")
    for k, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if s * 200 < ts and i * 200 < tinst:
            continue
        print(f"{k:28s} {100.0*s/ts:8.2f} {100.0*i/tinst:7.2f} {i/1e9:7.3f} {t/max(i,1):5.1f}")
    print(f"total: {ts} samples, {tinst/1e9:.3f} G warp instructions")


if __name__ == "__main__":
    main()
