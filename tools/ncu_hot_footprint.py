#!/usr/bin/env python3
"""Instruction-cache footprint of one kernel by source line: SASS instructions whose executed count is at least `frac` of the hottest
instruction's (the code the I-cache has to hold), from `ncu --page source --csv` joined with `nvdisasm -g` of the same build.

  python tools/ncu_hot_footprint.py <report.ncu-rep> <cubin> <mangled kernel name substring> [frac=0.01] [top=40]
"""
import collections, csv, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]
frac = float(sys.argv[4]) if len(sys.argv) > 4 else 0.01
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; ia, iexec = H.index("Address"), H.index("Instructions Executed")
inst = [(int(r[ia], 16), int(r[iexec] or 0)) for r in rows[hdr + 1:] if len(r) > iexec and r[0].startswith("0x")]
base = inst[0][0]; mx = max(e for _, e in inst)
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
line_of, cur = {}, ("?", 0)
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0])
for a, e in inst:
    k = line_of.get(a - base, ("?", 0)); agg[k][2] += 1
    if e >= frac * mx: agg[k][0] += 1; agg[k][1] += e
hot = sum(v[0] for v in agg.values())
print("kernel %s: %d SASS instructions, %d hot (executed >= %g of max) = %.1f KB" % (kname, len(inst), hot, frac, hot * 16 / 1024))
print("%-28s %6s %6s %14s" % ("file:line", "hot", "all", "executed"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s %6d %6d %14d" % ("%s:%d" % k, v[0], v[2], v[1]))
