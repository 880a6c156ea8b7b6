#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` (SASS) dump per kernel: executed warp-instructions and
stall samples by opcode, plus static code size."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kern = None
data = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kern = r[1][:70]
        data[kern] = []
        hdr = None
        continue
    if r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr and kern:
        data[kern].append(r)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
for k, rs in data.items():
    ex = collections.Counter()
    st = collections.Counter()
    tot_ex = tot_s = 0
    for r in rs:
        op = r[hdr["Source"]].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LD", "ST", "MUFU", "SHFL")) and "." in op else "")
        e = float(r[hdr["Instructions Executed"]] or 0)
        s = float(r[hdr["# Samples"]] or 0)
        ex[op] += e
        st[op] += s
        tot_ex += e
        tot_s += s
    print(f"== {k}\n   static SASS instrs {len(rs)} ({len(rs) * 16 / 1024:.0f} KB), executed {tot_ex / 1e6:.1f} M warp-instr, samples {tot_s:.0f}")
    for op, e in ex.most_common(top):
        print(f"   {op:14s} exec {100 * e / tot_ex:5.1f}%   samples {100 * st[op] / max(tot_s, 1):5.1f}%")
