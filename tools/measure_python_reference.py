#!/usr/bin/env python3
"""The Python reference itself on this machine's cores, the way config/parallel.py runs it (one `python -u abides.py -c <config> -s <seed>` process per
core, /root/reference/config/parallel.py:10-25) -- BASELINE.md section 3.  The reference cannot travel to the GPU box (nothing there may read
/root/reference), so this is measured in the build container and bench.py reports the recorded number with its provenance next to the GPU figures.

    python tools/measure_python_reference.py [n_procs] > profiles/r02_python_reference_cpu.json
"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
n = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tools", "shims") + ":" + REF)


def one(seed):
    d = tempfile.mkdtemp(prefix="abides_ref_")
    t0 = time.perf_counter()
    out = subprocess.run([sys.executable, "-u", os.path.join(REF, "abides.py"), "-c", "sparse_zi_1000", "-l", "m%d" % seed, "-s", str(seed)], cwd=d, env=env,
                         capture_output=True, text=True).stdout
    wall = time.perf_counter() - t0
    m = re.search(r"Event Queue elapsed: (\S+ days )?(\d+):(\d+):([\d.]+), messages: (\d+), messages per second: ([\d.]+)", out)
    if not m:
        return None
    el = int(m.group(2)) * 3600 + int(m.group(3)) * 60 + float(m.group(4))
    return dict(seed=seed, messages=int(m.group(5)), event_queue_elapsed_s=el, msgs_per_s=float(m.group(6)), process_wall_s=wall)


t0 = time.perf_counter()
with ThreadPoolExecutor(n) as ex:
    res = [r for r in ex.map(one, [1001 + i for i in range(n)]) if r]
wall = time.perf_counter() - t0
msgs = sum(r["messages"] for r in res)
cpu = open("/proc/cpuinfo").read()
model = re.search(r"model name\s*:\s*(.*)", cpu)
print(json.dumps({
    "what": "unmodified Python reference, config/parallel.py pattern: %d concurrent `python -u abides.py -c sparse_zi_1000 -s <seed>` processes" % n,
    "where": "build container (NOT the GPU box): %s, %d logical cores" % (model.group(1) if model else "?", os.cpu_count()),
    "processes": n, "cores": os.cpu_count(), "messages": msgs, "max_event_queue_elapsed_s": max(r["event_queue_elapsed_s"] for r in res),
    "msgs_per_s_event_loop": msgs / max(r["event_queue_elapsed_s"] for r in res), "sum_of_process_rates": sum(r["msgs_per_s"] for r in res),
    "msgs_per_s_wall": msgs / wall, "wall_s": wall, "runs": res}, indent=1))
