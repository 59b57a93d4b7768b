"""Shared test helpers: oracle -> device tape conversion, host-emulation harness loader."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libabx_host_emu.so")


def build_emu():
    """Compile tests/emu (the product's warp-uniform logic as plain C++; CPU CI only, never a fallback)."""
    src = os.path.join(EMU_DIR, "abx_host_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "marl_optimal_execution_b200", "csrc", f)
                    for f in ("abx_core.cuh", "abx_host_common.h")] + [os.path.join(ROOT, "include", "abides_b200.h")]
    if os.path.exists(EMU_LIB) and all(os.path.getmtime(EMU_LIB) >= os.path.getmtime(d) for d in deps):
        return EMU_LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-Wno-unknown-pragmas",
                           "-o", EMU_LIB, src])
    return EMU_LIB


def oracle_tapes(sims):
    """Device tape arrays (include/abides_b200.h: abx_sim_reset_tape) from finished OracleSim runs (one per env)."""
    bits, kinds, off, lat_to, lat_from = [], [], [0], [], []
    for s in sims:
        n = s.n_agents
        if s.variant == 3:                            # rmsc03: oracle stream order symbol, exchange, agents 1..n-1, kernel; global stream separate
            gk, gb = s.global_tape()                  # runtime draws; the oracle's __init__ megashock gap (drawn by the config) comes first
            g0 = s.global_exp_tape()[:1]
            glob = (np.concatenate([np.full(1, ord("e"), np.uint8), gk]), np.concatenate([g0.view(np.uint64), gb]))
            streams = [s.tape(0), s.tape(n + 1), (np.zeros(0, np.uint8), np.zeros(0, np.uint64)), glob]
            streams += [s.tape(a + 1) for a in range(1, n)]
            for k, b in streams:
                kinds.append(k)
                bits.append(b)
                off.append(off[-1] + len(b))
            info = np.array([s.agent_info(a) for a in range(n)])
            lat_to.append(info[:, 1].astype(np.float64))      # Noise/Value size (drawn by the config script)
            lat_from.append(info[:, 2].astype(np.float64))    # NoiseAgent.wakeup_time
            continue
        base = 4 if s.variant == 100 else 3           # oracle stream order: symbol, kernel, [latency], exchange, agents
        streams = [s.tape(0), s.tape(1)]
        streams.append(s.tape(2) if s.variant == 100 else (np.zeros(0, np.uint8), np.zeros(0, np.uint64)))
        g = s.global_exp_tape()
        streams.append((np.full(len(g), ord("e"), np.uint8), g.view(np.uint64)))
        for a in range(1, n):
            streams.append(s.tape(base + a - 1))
        for k, b in streams:
            kinds.append(k)
            bits.append(b)
            off.append(off[-1] + len(b))
        a, b = s.latency_vectors()
        lat_to.append(a)
        lat_from.append(b)
    return (np.concatenate(bits), np.concatenate(kinds), np.array(off, np.int64), np.concatenate(lat_to),
            np.concatenate(lat_from))
