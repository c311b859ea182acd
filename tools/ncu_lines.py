#!/usr/bin/env python
"""Executed instructions and warp-stall samples of one profiled kernel BY SOURCE LINE.

  python tools/ncu_lines.py <report.ncu-rep> <library.so> <kernel-symbol-substring> [top]

ncu's CSV export of the source page has no line column, so the SASS listing of the report (one row per
instruction, in address order) is zipped with `nvdisasm -g` of the SAME build of the library (its `//## File ...
line N` markers; built with -lineinfo).  Prints shares by file, by SASS opcode, and the `top` hottest lines."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, sym = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1])
    h, data = rows[1], rows[2:]
    ix = {k: i for i, k in enumerate(h)}
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, stdout=subprocess.DEVNULL, check=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], stdout=subprocess.PIPE,
                              stderr=subprocess.DEVNULL, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(sass) if l.startswith("//---") and sym in l][0]
    end = [i for i in range(start + 1, len(sass)) if sass[i].startswith("//---")][0]
    cur, inst = ("?", 0), []
    for l in sass[start:end]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            inst.append((m.group(2), cur))
    if len(inst) != len(data):
        sys.exit("the library is not the profiled build: %d SASS instructions against %d in the report" %
                 (len(inst), len(data)))
    by_line, by_file, by_op = (collections.defaultdict(lambda: [0, 0]) for _ in range(3))
    for r, (txt, (f, ln)) in zip(data, inst):
        ie, sm = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", txt)
        for d, k in ((by_line, (f, ln)), (by_file, os.path.basename(f)), (by_op, m.group(2) if m else "?")):
            d[k][0] += ie
            d[k][1] += sm
    ti = sum(v[0] for v in by_file.values())
    ts = sum(v[1] for v in by_file.values())
    print("warp instructions executed %d, warp-state samples %d" % (ti, ts))
    print("\nby file:")
    for k, v in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
        print("  %-30s inst %5.1f %%   samples %5.1f %%" % (k, 100.0 * v[0] / ti, 100.0 * v[1] / ts))
    print("\nby opcode:")
    for k, v in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:25]:
        print("  %-12s inst %5.1f %%   samples %5.1f %%" % (k, 100.0 * v[0] / ti, 100.0 * v[1] / ts))
    print("\nby source line:")
    src = {}
    for (f, ln), v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in src:
            try:
                src[f] = open(f).read().split("\n")
            except OSError:
                src[f] = []
        text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
        print("  inst %5.2f %%  samples %5.2f %%  %s:%d  %s" % (100.0 * v[0] / ti, 100.0 * v[1] / ts,
                                                                os.path.basename(f), ln, text))


if __name__ == "__main__":
    main()
