#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel: joins `ncu --page source --csv` (per-SASS-instruction samples and executed counts)
with `nvdisasm -g` line information of the same build (compile with -lineinfo).

  python tools/ncu_hot_lines.py gpurun_out/prof_run.ncu-rep <cubin built from the profiled sources> <mangled kernel name substring> [top]
"""
import collections, csv, re, subprocess, sys

rep, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ia, isamp, iexec = H.index("Address"), H.index("# Samples"), H.index("Instructions Executed")
inst = [(int(r[ia], 16), int(r[isamp] or 0), int(r[iexec] or 0), r[1].strip()) for r in rows[hdr + 1:] if len(r) > iexec and r[0].startswith("0x")]
base = inst[0][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
line_of, cur = {}, ("?", 0)
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_s = sum(i[1] for i in inst); tot_e = sum(i[2] for i in inst)
for a, s, e, _ in inst:
    k = line_of.get(a - base, ("?", 0))
    agg[k][0] += s; agg[k][1] += e; agg[k][2] += 1
print("kernel %s: %d SASS instructions, %d samples, %d warp-instructions executed" % (kname, len(inst), tot_s, tot_e))
print("%-28s %8s %7s %12s %7s %5s" % ("file:line", "samples", "%", "executed", "%", "SASS"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s %8d %6.2f%% %12d %6.2f%% %5d" % ("%s:%d" % k, v[0], 100.0 * v[0] / max(tot_s, 1), v[1], 100.0 * v[1] / max(tot_e, 1), v[2]))
