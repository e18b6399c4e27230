#!/usr/bin/env python
"""Attribute ncu per-instruction samples to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top=25]

ncu's CSV source page is per SASS instruction; this joins it, by instruction order, with
`nvdisasm -g` of the sm_100a cubin inside the in-tree library (built with -lineinfo) and sums
warp-stall samples and executed instructions per source line.
"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "aqsis_b200", "_lib", "libaqsis_b200_hider.so")


def disasm_lines(kernel_re):
    # NCU_CUBIN / NCU_SRC: a cubin + source of the commit the capture was taken at, when the tree has moved on
    cubin = os.environ.get("NCU_CUBIN")
    if not cubin:
        tmp = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
        cubin = [f for f in glob.glob(os.path.join(tmp, "*.cubin")) if os.path.basename(f).startswith("hider_kernels.")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    out, cur, line, on = {}, None, 0, False
    for l in txt.splitlines():
        m = re.match(r"\.text\.(\S+):", l)
        if m:
            cur = m.group(1)
            on = re.search(kernel_re, cur) is not None
            if on:
                out[cur] = []
            continue
        m = re.search(r'//## File ".*", line (\d+)', l)
        if m:
            line = int(m.group(1))
            continue
        if on and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            out[cur].append((line, l.strip()))
    return out


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                            capture_output=True, text=True).stdout
    rows = list(csv.reader(csvtxt.splitlines()))
    # several kernels/launches may follow each other: take the first block
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    kname = rows[hdr_i[0] - 1][1]
    hdr = rows[hdr_i[0]]
    end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
    body = rows[hdr_i[0] + 1:end]
    si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    l2i, shi = hdr.index("L2 Theoretical Sectors Global"), hdr.index("L1 Wavefronts Shared")
    stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    dis = disasm_lines(kre)
    fn = [k for k in dis if len(dis[k]) == len(body)]
    if not fn:
        print("instruction count mismatch", len(body), {k: len(v) for k, v in dis.items()})
        return
    lines = dis[fn[0]]
    src = open(os.environ.get("NCU_SRC") or os.path.join(ROOT, "aqsis_b200", "csrc", "hider_kernels.cu")).read().splitlines()
    agg = {}
    tot_s = tot_i = tot_l2 = tot_sh = 0
    stall_tot = {}
    for (ln, sass), r in zip(lines, body):
        s, i, t = int(r[si]), int(r[ii]), int(r[ti])
        a = agg.setdefault(ln, [0, 0, 0, 0, 0, {}])
        a[0] += s
        a[1] += i
        a[2] += t
        a[3] += int(r[l2i] or 0)
        a[4] += int(r[shi] or 0)
        for ci, nm in stall_cols:
            v = int(r[ci] or 0)
            if v:
                a[5][nm] = a[5].get(nm, 0) + v
                stall_tot[nm] = stall_tot.get(nm, 0) + v
        tot_s += s
        tot_i += i
        tot_l2 += int(r[l2i] or 0)
        tot_sh += int(r[shi] or 0)
    print(f"kernel {kname}: {len(body)} SASS instructions, {tot_s} samples, {tot_i} warp instructions, "
          f"{tot_l2} L2 sectors (global, theoretical), {tot_sh} shared wavefronts")
    print("stall samples: " + ", ".join(f"{k} {100.0 * v / max(tot_s, 1):.1f}%" for k, v in sorted(stall_tot.items(), key=lambda kv: -kv[1])[:8]))
    key = {"samples": 0, "inst": 1, "l2": 3, "shared": 4}[os.environ.get("NCU_SORT", "samples")]
    print(f"{'line':>5} {'samples%':>8} {'inst%':>7} {'lanes':>5} {'L2sec%':>6} {'shwf%':>6} {'top stall':>14}  source")
    for ln, (s, i, t, l2, sh, st) in sorted(agg.items(), key=lambda kv: -kv[1][key])[:top]:
        text = src[ln - 1].strip() if 0 < ln <= len(src) else "?"
        ts = max(st.items(), key=lambda kv: kv[1])[0] if st else "-"
        print(f"{ln:5d} {100.0 * s / max(tot_s, 1):8.2f} {100.0 * i / max(tot_i, 1):7.2f} {t / max(i, 1):5.1f} "
              f"{100.0 * l2 / max(tot_l2, 1):6.2f} {100.0 * sh / max(tot_sh, 1):6.2f} {ts:>14}  {text[:100]}")


if __name__ == "__main__":
    main()
