#!/usr/bin/env python3
"""Attribute an ncu SASS-level source page to CUDA source lines.

usage: sass_by_line.py <report.ncu-rep> <library.so> <kernel-substring> [top]

ncu --page source --csv lists per-SASS-instruction counters; nvdisasm -g gives the source line of
every SASS instruction of the same cubin (compiled with -lineinfo).  The two listings are joined by
instruction offset and summed per source line.  Runs on the CPU-only build box.
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
from pathlib import Path


def main():
    rep, lib, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.DictReader(lines[start:]))
    base = int(rows[0]["Address"], 16)
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", str(Path(lib).resolve())], cwd=td, capture_output=True)
        text = ""
        for cub in Path(td).glob("*.cubin"):
            t = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout
            if kern in t:
                text = t
                break
    sec = None
    cur = ("?", 0)
    line_of = {}
    for l in text.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            sec = m.group(1)
            continue
        if sec is None or kern not in sec:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    for r in rows:
        off = int(r["Address"], 16) - base
        key = line_of.get(off, ("?", 0))
        ie = int(r["Instructions Executed"] or 0)
        te = int(r["Thread Instructions Executed"] or 0)
        sm = int(r["# Samples"] or 0)
        is64 = 1 if re.search(r"\bD(FMA|MUL|ADD|SETP|MNMX)|MUFU\.\w+64", r["Source"]) else 0
        a = agg[key]
        a[0] += ie
        a[1] += te
        a[2] += sm
        a[3] += ie * is64
        tot[0] += ie
        tot[1] += te
        tot[2] += sm
    print(f"total warp-instr {tot[0]:,}  thread-instr {tot[1]:,}  avg lanes {tot[1] / max(tot[0], 1):.2f}  samples {tot[2]:,}")
    print(f"{'file:line':34s} {'warp-instr':>14s} {'%':>6s} {'fp64%':>6s} {'lanes':>6s} {'samples%':>8s}")
    for key, a in sorted(agg.items(), key=lambda t: -t[1][0])[:top]:
        print(f"{key[0] + ':' + str(key[1]):34s} {a[0]:14,d} {100 * a[0] / tot[0]:6.2f} {100 * a[3] / max(a[0], 1):6.1f} "
              f"{a[1] / max(a[0], 1):6.2f} {100 * a[2] / max(tot[2], 1):8.2f}")


if __name__ == "__main__":
    main()
