"""BatchedSim: thousands of independent ABIDES simulations stepped at once on one B200.

Python mirror of the reference's config + Kernel surface for the background-population configs
(config/sparse_zi_100.py, config/sparse_zi_1000.py -> Kernel.runner, Kernel.py:50-345), calling the
hand-written sm_100a kernels through the C ABI (include/abides_b200.h).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import EnvStats, SimConfig, TraceRec


def sparse_zi_config(variant, lib=None, **overrides):
    """abx_sim_config for config/sparse_zi_100.py (variant=100) or config/sparse_zi_1000.py (variant=1000)."""
    L = lib or _lib.load()
    cfg = SimConfig()
    _lib.check(L, L.abx_config_sparse_zi(int(variant), C.byref(cfg)), "abx_config_sparse_zi")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError("abx_sim_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg


def rmsc03_config(lib=None, pov_exec=False, **overrides):
    """abx_sim_config for config/rmsc03.py (50 Noise + 10 Value + 1 POV market maker + 2 Momentum agents, 09:30-09:45);
    pov_exec=True appends one POVExecutionAgent (agent/execution/baselines/pov_agent.py) as agent 64."""
    L = lib or _lib.load()
    cfg = SimConfig()
    if pov_exec:
        _lib.check(L, L.abx_config_rmsc03_pov(C.byref(cfg)), "abx_config_rmsc03_pov")
    else:
        _lib.check(L, L.abx_config_rmsc03(C.byref(cfg)), "abx_config_rmsc03")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError("abx_sim_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg


def rmsc01_config(lib=None, **overrides):
    """abx_sim_config for config/rmsc01.py (1 MarketMakerAgent + 50 ZI + 25 HBL + 24 Momentum agents, 09:30-16:00; the HBL agents read the
    exchange's order stream, agent/HeuristicBeliefLearningAgent.py:61-195)."""
    L = lib or _lib.load()
    cfg = SimConfig()
    _lib.check(L, L.abx_config_rmsc01(C.byref(cfg)), "abx_config_rmsc01")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError("abx_sim_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg


def rmsc02_config(lib=None, **overrides):
    """abx_sim_config for config/rmsc02.py: the rmsc01 population with the market maker and the momentum agents subscribed to MARKET_DATA
    (agent/ExchangeAgent.py:342-387), pairwise latency + noise, midnight-17:00."""
    L = lib or _lib.load()
    cfg = SimConfig()
    _lib.check(L, L.abx_config_rmsc02(C.byref(cfg)), "abx_config_rmsc02")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError("abx_sim_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg


def _num(x):
    """'{}'.format of a config literal: 1 -> '1', 0.8 -> '0.8' (config/sparse_zi_1000.py:211-219 writes eta as 1 or 0.8)."""
    return str(int(x)) if float(x) == int(x) else repr(float(x))


def agent_directory(cfg):
    """(names, types) of agents 1 .. n_agents-1 exactly as the config scripts name them: config/sparse_zi_1000.py:224-250
    ("ZI Agent {j} Type {k} [{R_min} <= R <= {R_max}, eta={eta}]" / "ZeroIntelligenceAgent Type ..."), config/rmsc03.py:118-200
    (NoiseAgent, Value Agent, POV_MARKET_MAKER_AGENT_, MOMENTUM_AGENT_), and the recorder's POV_EXECUTION_AGENT / ExecutionAgent."""
    names, types = [], []
    j = 1
    if cfg.population == 0:
        for k in range(cfg.n_groups):
            g = cfg.groups[k]
            strat = "Type {} [{} <= R <= {}, eta={}]".format(k + 1, g.r_min, g.r_max, _num(g.eta))
            for _ in range(g.count):
                names.append("ZI Agent {} {}".format(j, strat)); types.append("ZeroIntelligenceAgent {}".format(strat)); j += 1
    elif cfg.population == 3:                         # config/rmsc01.py:98-210
        for count, name, typ in ((cfg.n_mm_agents, "MARKET_MAKER_AGENT_{}", "MarketMakerAgent"), (cfg.groups[0].count, "ZI_AGENT_{}", "ZeroIntelligenceAgent"),
                                 (cfg.groups[1].count, "HBL_AGENT_{}", "HeuristicBeliefLearningAgent"), (cfg.n_momentum_agents, "MOMENTUM_AGENT_{}", "MomentumAgent")):
            for _ in range(count):
                names.append(name.format(j)); types.append(typ); j += 1
    else:
        for count, name, typ in ((cfg.n_noise_agents, "NoiseAgent {}", "NoiseAgent"), (cfg.n_value_agents, "Value Agent {}", "ValueAgent"),
                                 (cfg.n_mm_agents, "POV_MARKET_MAKER_AGENT_{}", "POVMarketMakerAgent"), (cfg.n_momentum_agents, "MOMENTUM_AGENT_{}", "MomentumAgent")):
            for _ in range(count):
                names.append(name.format(j)); types.append(typ); j += 1
        for _ in range(cfg.n_pov_exec):
            names.append("POV_EXECUTION_AGENT"); types.append("ExecutionAgent"); j += 1
    return names, types


def format_kernel_summary(names, types, holdings, messages, starting_cash, symbol="JPM", elapsed_s=None):
    """The text Kernel.runner leaves on stdout for one simulation, from the per-agent rows
    (id, shares, cash, marked to market, surplus) of abx_sim_holdings:
      * per trading agent "Final holdings for {name}: { SYM: n, CASH: c }.  Marked to market: m" (agent/TradingAgent.py:115-126; fmtHoldings
        :670-680 -- a flat position has no symbol entry, orderExecuted deletes it);
      * "Event Queue elapsed: {Timedelta}, messages: {N}, messages per second: {R:0.1f}" (Kernel.py:321-327);
      * "Mean ending value by agent type:" + "{type}: {int(round(mean gain))}" in first-seen order (Kernel.py:337-341, TradingAgent.py:130-138);
      * "Simulation ending!" (Kernel.py:343)."""
    lines, gain, count = [], {}, {}
    for (aid, shares, cash, mtm, _), name, typ in zip(holdings, names, types):
        h = ("{}: {}, ".format(symbol, int(shares)) if int(shares) != 0 else "") + "CASH: {}".format(int(cash))
        lines.append("Final holdings for {}: {}.  Marked to market: {}".format(name, "{ " + h + " }", int(mtm)))
        gain[typ] = gain.get(typ, 0) + int(mtm) - int(starting_cash)
        count[typ] = count.get(typ, 0) + 1
    if elapsed_s is not None:
        us = int(round(elapsed_s * 1e6))
        d, rem = divmod(us, 86400 * 10 ** 6)
        hh, rem = divmod(rem, 3600 * 10 ** 6)
        mm, rem = divmod(rem, 60 * 10 ** 6)
        ss, frac = divmod(rem, 10 ** 6)
        td = "{} days {:02d}:{:02d}:{:02d}".format(d, hh, mm, ss) + (".{:06d}".format(frac) if frac else "")      # str(pd.Timedelta)
        lines.append("Event Queue elapsed: {}, messages: {}, messages per second: {:0.1f}".format(td, int(messages), int(messages) / max(elapsed_s, 1e-12)))
    lines.append("Mean ending value by agent type:")
    for typ in gain:
        lines.append("{}: {:d}".format(typ, int(round(gain[typ] / count[typ]))))
    lines.append("Simulation ending!")
    return lines


class BatchedSim:
    """n_envs independent simulations of one population config on one GPU.

    run(until_ns) advances every environment's event loop (Kernel.py:190-292) to `until_ns`;
    stats() returns the per-environment counters (messages == the reference's ttl_messages).
    """

    def __init__(self, cfg, n_envs, device=0, lib_path=None):
        self._L = _lib.load(lib_path)
        self.cfg = cfg
        self.n_envs = int(n_envs)
        self.n_agents = int(cfg.n_agents)
        self._h = C.c_void_p()
        _lib.check(self._L, self._L.abx_sim_create(C.byref(cfg), self.n_envs, int(device), C.byref(self._h)),
                   "abx_sim_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.abx_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reset ----
    def reset(self, seeds, stream=None):
        """Philox mode: environment e draws every random variate from counter streams keyed by seeds[e]."""
        seeds = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, dtype=np.uint64), (self.n_envs,)))
        _lib.check(self._L, self._L.abx_sim_reset_philox(self._h, seeds.ctypes.data_as(C.POINTER(C.c_uint64)), stream),
                   "abx_sim_reset_philox")

    def reset_tape(self, tape_bits, tape_kinds, tape_offsets, lat_to_exchange, lat_from_exchange, stream=None):
        """Tape mode: replay the RandomState draws of a recorded reference run (include/abides_b200.h)."""
        n_streams = self.n_agents + 3
        bits = np.ascontiguousarray(tape_bits, dtype=np.uint64)
        kinds = np.ascontiguousarray(tape_kinds, dtype=np.uint8)
        off = np.ascontiguousarray(tape_offsets, dtype=np.int64)
        lt = np.ascontiguousarray(lat_to_exchange, dtype=np.float64)
        lf = np.ascontiguousarray(lat_from_exchange, dtype=np.float64)
        if off.shape != (self.n_envs * n_streams + 1,) or off[-1] != bits.size or kinds.size != bits.size:
            raise ValueError("tape arrays have inconsistent shapes")
        if lt.size != self.n_envs * self.n_agents or lf.size != lt.size:
            raise ValueError("latency vectors must be [n_envs * n_agents]")
        _lib.check(self._L, self._L.abx_sim_reset_tape(
            self._h, bits.ctypes.data_as(C.POINTER(C.c_uint64)), kinds.ctypes.data_as(C.POINTER(C.c_uint8)),
            off.ctypes.data_as(C.POINTER(C.c_int64)), lt.ctypes.data_as(C.POINTER(C.c_double)),
            lf.ctypes.data_as(C.POINTER(C.c_double)), stream), "abx_sim_reset_tape")

    def reset_tape_shared(self, n_tapes, tape_bits, tape_kinds, tape_offsets, lat_to_exchange, lat_from_exchange, stream=None):
        """Tape mode with `n_tapes` recorded runs shared by all environments: environment e replays run e % n_tapes (arrays as reset_tape, for
        n_tapes runs)."""
        n_streams = self.n_agents + 3
        bits = np.ascontiguousarray(tape_bits, dtype=np.uint64)
        kinds = np.ascontiguousarray(tape_kinds, dtype=np.uint8)
        off = np.ascontiguousarray(tape_offsets, dtype=np.int64)
        lt = np.ascontiguousarray(lat_to_exchange, dtype=np.float64)
        lf = np.ascontiguousarray(lat_from_exchange, dtype=np.float64)
        if off.shape != (n_tapes * n_streams + 1,) or off[-1] != bits.size or kinds.size != bits.size or lt.size != n_tapes * self.n_agents or lf.size != lt.size:
            raise ValueError("tape arrays have inconsistent shapes")
        _lib.check(self._L, self._L.abx_sim_reset_tape_shared(
            self._h, int(n_tapes), bits.ctypes.data_as(C.POINTER(C.c_uint64)), kinds.ctypes.data_as(C.POINTER(C.c_uint8)),
            off.ctypes.data_as(C.POINTER(C.c_int64)), lt.ctypes.data_as(C.POINTER(C.c_double)), lf.ctypes.data_as(C.POINTER(C.c_double)), stream),
            "abx_sim_reset_tape_shared")

    # ---- stepping ----
    def run(self, until_ns=None, stream=None):
        until = int(self.cfg.stop_ns) + 10 ** 15 if until_ns is None else int(until_ns)
        _lib.check(self._L, self._L.abx_sim_run(self._h, until, stream), "abx_sim_run")

    def run_each(self, until_ns_host_ptr, stream=None):
        """Per-environment horizons: `until_ns_host_ptr` is the address of a HOST int64[n_envs] array (pinned
        memory, e.g. a torch pinned tensor's data_ptr(), keeps the copy asynchronous)."""
        _lib.check(self._L, self._L.abx_sim_run_each(self._h, C.c_void_p(until_ns_host_ptr), stream), "abx_sim_run_each")

    def finalize(self, stream=None):
        _lib.check(self._L, self._L.abx_sim_finalize(self._h, stream), "abx_sim_finalize")

    # ---- results ----
    def stats(self, stream=None, out=None):
        """Per-environment abx_env_stats as a numpy structured array (device -> host copy, synchronises `stream`).
        `out` may be a preallocated (e.g. pinned) buffer exposing the array interface with 112 bytes per env."""
        if out is None:
            out = np.zeros(self.n_envs, dtype=_lib.STATS_DTYPE)
        _lib.check(self._L, self._L.abx_sim_stats(self._h, out.ctypes.data, stream), "abx_sim_stats")
        return out

    def stats_into(self, device_ptr, stream=None):
        """Leave the abx_env_stats records on the device (device_ptr: e.g. torch tensor .data_ptr(), 112 B/env)."""
        _lib.check(self._L, self._L.abx_sim_stats_device(self._h, C.c_void_p(device_ptr), stream), "abx_sim_stats_device")

    def pov_exec(self, env, stream=None):
        """POVExecutionAgent of one environment: (remaining quantity, executed orders, open orders)."""
        out = np.zeros(3, dtype=np.int64)
        _lib.check(self._L, self._L.abx_sim_pov_exec(self._h, int(env), out.ctypes.data_as(C.POINTER(C.c_int64)), stream), "abx_sim_pov_exec")
        return out

    def holdings(self, env, stream=None):
        out = np.zeros((self.n_agents - 1, 5), dtype=np.int64)
        _lib.check(self._L, self._L.abx_sim_holdings(self._h, int(env), out.ctypes.data_as(C.POINTER(C.c_int64)), stream),
                   "abx_sim_holdings")
        return out

    def kernel_summary(self, env=0, symbol="JPM", elapsed_s=None, stream=None):
        """Kernel.runner's stdout report of environment `env` (call after run() + finalize()): see format_kernel_summary."""
        names, types = agent_directory(self.cfg)
        st = self.stats(stream=stream)[int(env)]
        return format_kernel_summary(names, types, self.holdings(env, stream=stream), int(st["messages"]), int(self.cfg.starting_cash), symbol, elapsed_s)

    def book_snapshot(self, env, is_bid, depth, stream=None):
        out = np.zeros(2 * max(depth, 1), dtype=np.int32)
        n = C.c_int32(0)
        _lib.check(self._L, self._L.abx_sim_book_snapshot(self._h, int(env), int(bool(is_bid)), int(depth),
                                                          out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n), stream),
                   "abx_sim_book_snapshot")
        return [(int(out[2 * i]), int(out[2 * i + 1])) for i in range(n.value)]

    def trace(self, env, stream=None):
        cap = int(self.cfg.trace_cap)
        out = np.zeros(max(cap, 1), dtype=_lib.TRACE_DTYPE)
        n = C.c_int32(0)
        _lib.check(self._L, self._L.abx_sim_trace(self._h, int(env), out.ctypes.data, cap, C.byref(n), stream),
                   "abx_sim_trace")
        return out[: n.value]

    def draw_log(self, env, stream=None):
        """Parity instrumentation (cfg.draw_log_cap > 0, Philox mode): every standard variate environment `env` has drawn so far, in draw
        order, as (stream, kind, bits) arrays -- stream numbering of reset_tape (0 symbol, 1 kernel, 2 latency model, 3 global, 3 + a agent a)."""
        cap = int(self.cfg.draw_log_cap)
        raw = np.zeros(max(cap, 1) * 4, dtype=np.uint32)
        n = C.c_int32(0)
        _lib.check(self._L, self._L.abx_sim_draw_log(self._h, int(env), raw.ctypes.data, cap, C.byref(n), stream), "abx_sim_draw_log")
        r = raw[: 4 * n.value].reshape(-1, 4)
        return (r[:, 0] & 0xFFFFFF).astype(np.int64), (r[:, 0] >> 24).astype(np.uint8), r[:, 1].astype(np.uint64) | (r[:, 2].astype(np.uint64) << np.uint64(32))

    def draw_tapes(self, env, stream=None):
        """The draw log split per stream, in the layout reset_tape / the oracle's external tapes take: (bits, kinds, offsets[n_agents + 4])."""
        st, kinds, bits = self.draw_log(env, stream)
        order = np.argsort(st, kind="stable")
        counts = np.bincount(st, minlength=self.n_agents + 3)[: self.n_agents + 3]
        off = np.zeros(self.n_agents + 4, dtype=np.int64)
        off[1:] = np.cumsum(counts)
        return bits[order], kinds[order], off

    def agent_init(self, env, stream=None):
        """Start-of-run trader state as the reset drew it: dict(theta [n_agents, 20], lat_to, lat_from, sizes, wakes)."""
        n = self.n_agents
        out = dict(theta=np.zeros((n, 20), np.int32), lat_to=np.zeros(n), lat_from=np.zeros(n), sizes=np.zeros(n, np.int32), wakes=np.zeros(n, np.int64))
        _lib.check(self._L, self._L.abx_sim_agent_init(self._h, int(env), out["theta"].ctypes.data, out["lat_to"].ctypes.data, out["lat_from"].ctypes.data,
                                                       out["sizes"].ctypes.data, out["wakes"].ctypes.data, stream), "abx_sim_agent_init")
        return out

    def split_trace(self, env):
        """Trace -> (pops[n,5], notes[n,13], snaps[n,16]) int64 arrays in tools/record_reference.py row layout."""
        tr = self.trace(env)
        pops = tr[tr["tag"] == 0]
        notes = tr[tr["tag"] == 1]
        snaps = tr[tr["tag"] == 2]
        p = np.zeros((len(pops), 5), np.int64)
        p[:, 0], p[:, 1] = pops["t"], pops["a"]
        p[:, 2:5] = pops["v"][:, 0:3]
        nt = np.zeros((len(notes), 13), np.int64)
        nt[:, 0], nt[:, 1] = notes["t"], notes["a"]
        nt[:, 2:13] = notes["v"][:, 0:11]
        sn = snaps["v"].astype(np.int64)
        return p, nt, sn

    @property
    def launch_count(self):
        return int(self._L.abx_sim_launch_count(self._h))

    @property
    def device_bytes(self):
        return int(self._L.abx_sim_device_bytes(self._h))
