#!/usr/bin/env python3
"""Estimate the instruction-cache footprint of a NEW build without a GPU: source lines that were hot in a profiled build (ncu source page
of that build + its cubin) are looked up in the new cubin's line table; reports hot instructions and 128-byte lines touched.

  python tools/static_hot_footprint.py <report.ncu-rep> <profiled cubin> <new cubin> <mangled kernel name substring> [frac=0.01]
"""
import collections, csv, re, subprocess, sys
rep, cub0, cub1, kname = sys.argv[1:5]
frac = float(sys.argv[5]) if len(sys.argv) > 5 else 0.01

def line_table(cubin):
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
    out, cur = [], ("?", 0)
    for l in sass[start + 1:]:
        if l.startswith(".text.") or l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", l)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2)))
    return out

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; ia, iexec = H.index("Address"), H.index("Instructions Executed")
inst = [(int(r[ia], 16), int(r[iexec] or 0)) for r in rows[hdr + 1:] if len(r) > iexec and r[0].startswith("0x")]
base = inst[0][0]; mx = max(e for _, e in inst)
t0 = {a: k for a, k, _ in line_table(cub0)}
hot_lines = collections.Counter()
for a, e in inst:
    k = t0.get(a - base)
    if k: hot_lines[k] = max(hot_lines[k], e)
hot = {k for k, e in hot_lines.items() if e >= frac * mx}
for name, cub in (("profiled", cub0), ("new", cub1)):
    t = line_table(cub)
    h = [(a, k) for a, k, _ in t if k in hot]
    print("%-9s %5d instructions, %5d on hot source lines (%.1f KB), %4d 128-byte lines (%.1f KB)" % (name, len(t), len(h), len(h) * 16 / 1024, len({a // 128 for a, _ in h}), len({a // 128 for a, _ in h}) * 128 / 1024))
