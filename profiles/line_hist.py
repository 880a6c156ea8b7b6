#!/usr/bin/env python3
"""Attribute executed warp-instructions / stall samples of one kernel to source lines.

  nvdisasm -g -c <cubin>  gives the SASS with `//## File "...", line N` markers (compile with -lineinfo);
  ncu --page source --csv gives per-SASS-instruction counters in the same order.
Usage: line_hist.py <ncu_source.csv> <kernel substring> <libspecloss.so> [top]
Prints per-file totals, then the hottest source lines, then totals per marked region
(`// [region: name]` comments in specloss_kernels.cuh start a region that lasts until the next marker)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

src_csv, kern_sub, so_path = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
sec_override = sys.argv[5] if len(sys.argv) > 5 else None          # mangled-name substring of the .text section

rows = list(csv.reader(open(src_csv)))
kern, data, hdr = None, {}, None
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        kern = r[1]
        data[kern] = []
        hdr = None
    elif r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr is not None and kern is not None:
        data[kern].append(r)
match = [k for k in data if kern_sub in k.replace("(int)", "").replace("(bool)", "").replace(" ", "")]
assert match, (kern_sub, list(data))
kname = match[0]
prof = data[kname]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# mangled-name match: pick the .text section whose demangled template args match
m = re.search(r"transform_kernel<(\d+),(\d+),(\d+),(\d+)>", kname.replace("(int)", "").replace("(bool)", "").replace(" ", ""))
if m:
    n, kind, grad, win = m.groups()
    sec = f"transform_kernelILi{n}ELi{kind}ELb{grad}ELi{win}EE"
else:
    sec = kern_sub
if sec_override:
    sec = sec_override
lines = sass.split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and sec in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//-----")), len(lines))
cur = ("?", 0)
instr_lines = []
for l in lines[start:end]:
    mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if mm:
        cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        instr_lines.append(cur)
assert len(instr_lines) == len(prof), (len(instr_lines), len(prof))

# regions from markers in the kernel source
regions = []
ksrc = os.path.join(os.path.dirname(os.path.abspath(so_path)), "csrc", "specloss_kernels.cuh")
for i, l in enumerate(open(ksrc), 1):
    mm = re.search(r"\[region: ([^\]]+)\]", l)
    if mm:
        regions.append((i, mm.group(1)))


def region_of(f, ln):
    if f != "specloss_kernels.cuh":
        return None
    name = "(before first marker)"
    for start_ln, nm in regions:
        if ln >= start_ln:
            name = nm
    return name


ex_i, st_i = hdr["Instructions Executed"], hdr["# Samples"]
by_file, by_line, by_region = collections.Counter(), collections.Counter(), collections.Counter()
s_file, s_line, s_region = collections.Counter(), collections.Counter(), collections.Counter()
tot = tots = 0.0
last_region = "(before first marker)"
for (f, ln), r in zip(instr_lines, prof):
    e, s = float(r[ex_i] or 0), float(r[st_i] or 0)
    reg = region_of(f, ln)
    if reg is None:
        reg = last_region + " / " + f      # inlined helper (codelets, intrinsics): charge to the enclosing region
    else:
        last_region = reg
    by_file[f] += e; by_line[(f, ln)] += e; by_region[reg] += e
    s_file[f] += s; s_line[(f, ln)] += s; s_region[reg] += s
    tot += e; tots += s
print(f"kernel {kname[:90]}\n  {len(prof)} SASS instrs, {tot / 1e6:.2f} M executed warp-instrs, {tots:.0f} samples")
print("-- by file")
for f, e in by_file.most_common():
    print(f"  {100 * e / tot:5.1f}% exec  {100 * s_file[f] / max(tots, 1):5.1f}% samples  {f}")
print("-- by region")
for reg, e in sorted(by_region.items(), key=lambda t: -t[1]):
    print(f"  {100 * e / tot:5.1f}% exec  {100 * s_region[reg] / max(tots, 1):5.1f}% samples  {reg}")
print(f"-- top {top} lines")
for (f, ln), e in by_line.most_common(top):
    print(f"  {100 * e / tot:5.1f}% exec  {100 * s_line[(f, ln)] / max(tots, 1):5.1f}% samples  {f}:{ln}")
print(f"-- top {top} lines by stall samples")
for (f, ln), sm in s_line.most_common(top):
    print(f"  {100 * sm / max(tots, 1):5.1f}% samples  {100 * by_line[(f, ln)] / tot:5.1f}% exec  {f}:{ln}")
