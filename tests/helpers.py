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
        if s.variant in (3, 1):                       # rmsc03 / rmsc01: oracle stream order symbol, exchange, agents 1..n-1, kernel; global stream separate
            gk, gb = s.global_tape()                  # runtime draws; the oracle's __init__ megashock gap (drawn by the config) comes first
            g0 = s.global_exp_tape()[:1]
            glob = (np.concatenate([np.full(1, ord("e"), np.uint8), gk]), np.concatenate([g0.view(np.uint64), gb]))
            streams = [s.tape(0), s.tape(n + 1), (np.zeros(0, np.uint8), np.zeros(0, np.uint64)), glob]
            streams += [s.tape(a + 1) for a in range(1, n)]
            for k, b in streams:
                kinds.append(k)
                bits.append(b)
                off.append(off[-1] + len(b))
            if s.variant == 1:                        # rmsc01 / rmsc02: every agent draws its own parameters from its own stream; latency matrix row / column 0
                a, b = s.latency_vectors()
                lat_to.append(a)
                lat_from.append(b)
                continue
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


def oracle_rmsc01_config(stop_ns=(9 * 3600 + 45 * 60) * 10 ** 9):
    """config/rmsc01.py as the oracle states it (abo_default_config(1): 1 MarketMakerAgent + 50 ZI + 25 HBL + 24 Momentum agents, zero latency), cut at
    `stop_ns`: the full day is ~2 M messages under the current reference code (tests/rmsc01.txt's 128 918 predates it), so recordings and
    parity runs stop early."""
    import ctypes as C
    from marl_optimal_execution_b200 import _lib
    from oracle.oracle import lib
    cfg = _lib.SimConfig()
    assert lib().abo_default_config(1, C.addressof(cfg)) == 0
    cfg.stop_ns = stop_ns
    return cfg


def oracle_rmsc02_config(stop_ns=17 * 3600 * 10 ** 9):
    """config/rmsc02.py as the oracle states it (abo_default_config(2)): the rmsc01 population with the market maker and the momentum agents in subscription
    mode, the sparse_zi_1000 latency model (uniform 21 us .. 13 ms matrix, 6-entry noise), midnight .. 17:00."""
    import ctypes as C
    from marl_optimal_execution_b200 import _lib
    from oracle.oracle import lib
    cfg = _lib.SimConfig()
    assert lib().abo_default_config(2, C.addressof(cfg)) == 0
    cfg.stop_ns = stop_ns
    return cfg


def oracle_rmsc02(seed, stop_ns, trace):
    from oracle.oracle import OracleSim
    o = OracleSim.from_config(oracle_rmsc02_config(stop_ns), seed, trace)
    return o, o.run()


def oracle_rmsc01(seed, stop_ns, trace):
    """A finished OracleSim of the rmsc01 population and its message count."""
    from oracle.oracle import OracleSim
    o = OracleSim.from_config(oracle_rmsc01_config(stop_ns), seed, trace)
    return o, o.run()


def oracle_rerun_of_philox_env(sim, env, init, trace):
    """Re-run environment `env` of a Philox-seeded BatchedSim (cfg.draw_log_cap > 0) through the oracle: the oracle is built from the SAME
    abx_sim_config, every RandomState is replaced by the standard variates the environment drew (sim.draw_tapes) and the start-of-run state
    (`init` = sim.agent_init(env), taken right after reset) is installed.  Returns the finished OracleSim and its message count."""
    from oracle.oracle import OracleSim
    bits, kinds, off = sim.draw_tapes(env)
    o = OracleSim.from_config(sim.cfg, 0, trace)
    o.set_external(bits, kinds, off, theta=init["theta"], lat_to=init["lat_to"], lat_from=init["lat_from"], sizes=init["sizes"], wakes=init["wakes"])
    n = o.run()
    assert o.rng_error() == 0, "external tape underrun / kind mismatch: %d" % o.rng_error()
    assert o.external_unread() == 0, "the oracle left %d of the environment's draws unread" % o.external_unread()
    return o, n


def assert_env_equals_oracle(sim, e, o, n, st, traces=True, holdings_cols=5, hashed=True):
    """Counters, L1, fundamental, conservation sums, holdings (and pop hash / full traces when the run was instrumented) of environment e == oracle."""
    from marl_optimal_execution_b200 import _lib
    assert int(st["messages"][e]) == n, (int(st["messages"][e]), n)
    assert int(st["flags"][e]) == _lib.F_DONE, hex(int(st["flags"][e]))
    if hashed:
        assert int(st["pop_hash"][e]) == o.pop_hash()
    if traces:
        p, nt, sn = sim.split_trace(e)
        for name, a, b in (("pops", p, o.trace("pops")), ("notes", nt, o.trace("notes")), ("snaps", sn, o.trace("snaps"))):
            assert a.shape == b.shape, (name, a.shape, b.shape)
            d = np.nonzero((a != b).any(axis=1))[0]
            assert len(d) == 0, (name, int(d[0]), a[d[0]], b[d[0]])
    assert np.array_equal(sim.holdings(e)[:, :holdings_cols], o.holdings()[:, :holdings_cols])
    for f, c in (("limit_orders", "limit"), ("cancels", "cancel"), ("fills", "fills"), ("spread_queries", "spread_queries"), ("max_queue", "max_queue"),
                 ("uniq", "uniq"), ("orders_allocated", "orders_allocated")):
        assert int(st[f][e]) == o.counter(c), (f, int(st[f][e]), o.counter(c))
    l1 = o.book_l1()
    assert (int(st["best_bid"][e]), int(st["best_bid_qty"][e]), int(st["best_ask"][e]), int(st["best_ask_qty"][e]), int(st["last_trade"][e])) == tuple(int(x) for x in l1)
    assert int(st["fundamental"][e]) == o.fundamental()


def oracle_events(o):
    """The exchange event log of a finished OracleSim run (trace flags OPS | NOTES | SNAPS) in the layout of the device event ring: chronological
    torch tensors (t, kind, a, b, valid) with one row -- order arrival, BEST_BID / BEST_ASK (util/OrderBook.py:114-128) and LAST_TRADE (:131-141)
    per handleLimitOrder call, exactly what marl_optimal_execution_b200.realism reduces."""
    import torch
    ops, notes, snaps = o.trace("ops"), o.trace("notes"), o.trace("snaps")
    rows = []
    ex = notes[notes[:, 2] == 8]                      # ORDER_EXECUTED rows come in pairs (incoming copy first)
    inc = ex[0::2]
    k = 0
    for i in range(len(ops)):
        if ops[i, 1] != 0:
            continue
        t, oid = int(ops[i, 0]), int(ops[i, 3])
        rows.append((t, 0, int(ops[i, 5]), int(ops[i, 6]) if ops[i, 4] else -int(ops[i, 6])))
        sn = snaps[i]
        if sn[0] > 0:
            rows.append((t, 1, int(sn[3]), int(sn[4])))
        if sn[1] > 0:
            rows.append((t, 2, int(sn[9]), int(sn[10])))
        q = pq = 0
        while k < len(inc) and inc[k, 0] == t and inc[k, 3] == oid:
            q += int(inc[k, 5]); pq += int(inc[k, 5]) * int(inc[k, 7]); k += 1
        if q:
            rows.append((t, 3, int(round(pq / q)), q))
    r = np.array(rows, dtype=np.int64)
    tt = torch.from_numpy(r.T.copy())
    return tt[0][None], tt[1][None], tt[2][None], tt[3][None], torch.ones(1, len(r), dtype=torch.bool)
