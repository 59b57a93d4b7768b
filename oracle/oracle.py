"""ctypes wrapper around oracle/libabides_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module (see oracle/abides_oracle.h).  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libabides_oracle.so")

TRACE_POPS, TRACE_OPS, TRACE_NOTES, TRACE_SNAPS, TRACE_TAPES = 1, 2, 4, 8, 16
TRACE_ALL = 31


def build(force=False):
    """Compile the C restatement (gcc, -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("abides_oracle.c", "abides_oracle.h", os.path.join("..", "include", "abides_b200.h"))]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libabides_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    vp, i64, i32, u32, u64, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_uint64, C.c_double
    P = C.POINTER

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("abo_rng_new", vp, u32)
    sig("abo_rng_free", None, vp)
    sig("abo_rng_u32", u32, vp)
    sig("abo_rng_double", dbl, vp)
    sig("abo_rng_gauss", dbl, vp)
    sig("abo_rng_std_exponential", dbl, vp)
    sig("abo_rng_randint", i64, vp, i64, i64)
    sig("abo_book_new", vp, i32)
    sig("abo_book_free", None, vp)
    sig("abo_book_set_time", None, vp, i64)
    sig("abo_book_limit", None, vp, i64, i64, i32, i64, i64)
    sig("abo_book_cancel", None, vp, i64, i64, i32, i64)
    sig("abo_book_modify", None, vp, i64, i64, i32, i64, i64, i64, i64)
    sig("abo_book_inside", i32, vp, i32, i32, P(i64))
    sig("abo_book_last_trade", i64, vp)
    sig("abo_book_transacted_volume", i64, vp, i64)
    sig("abo_book_n_levels", i32, vp, i32)
    sig("abo_book_n_resting", i32, vp)
    sig("abo_book_n_notes", i64, vp)
    sig("abo_book_notes", P(i64), vp)
    sig("abo_book_clear_notes", None, vp)
    sig("abo_book_level_orders", i32, vp, i32, i32, P(i64), i32)
    sig("abo_sim_new_sparse_zi", vp, i32, u32, i32)
    sig("abo_sim_new_rmsc03", vp, u32, i32)
    sig("abo_sim_new_rmsc03_pov", vp, u32, i32, dbl, i64, i32, i64, i64, i64, i64)
    sig("abo_sim_pov_exec", None, vp, P(i64))
    sig("abo_default_config", i32, i32, vp)
    sig("abo_sim_new_config", vp, vp, u32, i32)
    sig("abo_sim_set_external", i32, vp, P(u64), P(C.c_uint8), P(i64), P(C.c_int32), P(dbl), P(dbl), P(C.c_int32), P(i64))
    sig("abo_sim_rng_error", i32, vp)
    sig("abo_sim_external_unread", i64, vp)
    sig("abo_sim_global_tape", i64, vp, P(P(C.c_uint8)), P(P(u64)))
    sig("abo_sim_agent_info", None, vp, i32, P(i64))
    sig("abo_sim_free", None, vp)
    sig("abo_sim_run", i64, vp)
    sig("abo_sim_run_until", i64, vp, i64, P(i32))
    sig("abo_sim_start", None, vp)
    sig("abo_sim_stop", None, vp)
    sig("abo_sim_n_agents", i32, vp)
    sig("abo_sim_n_pops", i64, vp)
    sig("abo_sim_holdings", None, vp, P(i64))
    sig("abo_sim_pop_hash", u64, vp)
    sig("abo_sim_n_hash_ckpt", i64, vp)
    sig("abo_sim_hash_ckpt", P(u64), vp)
    sig("abo_sim_note_hash", u64, vp)
    sig("abo_sim_snap_hash", u64, vp)
    sig("abo_sim_trace", i64, vp, i32, P(P(i64)))
    sig("abo_sim_n_streams", i32, vp)
    sig("abo_sim_tape", i64, vp, i32, P(P(C.c_uint8)), P(P(u64)))
    sig("abo_sim_stream_seed", u32, vp, i32)
    sig("abo_sim_global_exp_tape", i64, vp, P(P(dbl)))
    sig("abo_sim_theta", None, vp, i32, P(C.c_int32))
    sig("abo_sim_latency_vectors", None, vp, P(dbl), P(dbl))
    sig("abo_sim_zi_params", None, vp, i32, P(dbl))
    sig("abo_sim_counter", i64, vp, i32)
    sig("abo_sim_book_l1", None, vp, P(i64))
    sig("abo_sim_fundamental", i64, vp)
    sig("abo_env_new", vp, P(i64), i64, dbl, i32, i32)
    sig("abo_env_new2", vp, P(i64), i64, dbl, i32, i32, i64)
    sig("abo_env_free", None, vp)
    sig("abo_env_step", i32, vp, P(dbl), P(dbl), P(i32))
    sig("abo_env_n_pops", i64, vp)
    sig("abo_env_pop_hash", u64, vp)
    sig("abo_env_note_hash", u64, vp)
    sig("abo_env_snap_hash", u64, vp)
    sig("abo_env_n_hash_ckpt", i64, vp)
    sig("abo_env_hash_ckpt", P(u64), vp)
    sig("abo_env_trace", i64, vp, i32, P(P(i64)))
    sig("abo_env_final", None, vp, P(dbl))
    sig("abo_env_counter", i64, vp, i32)
    sig("abo_dq_set_schedule", i32, vp, i32, P(i64), i32)
    sig("abo_dq_new", vp, P(i64), i64, i32, P(i64), i32, i32, i32, i64, i64, i64, i32, i64, i32)
    sig("abo_dq_step", i32, vp, i32, P(dbl), P(dbl), P(dbl), P(i32))
    sig("abo_dq_error", i32, vp)
    sig("abo_dq_series", i64, vp, i32, P(P(dbl)))
    sig("abo_dq_holdings", None, vp, P(i64))
    sig("abo_dq_exec_final", None, vp, i32, P(dbl))
    _lib = L
    return L


class NumpyLegacyRng:
    """MT19937 + legacy numpy distributions (restated in C); checked against numpy itself in tests."""

    def __init__(self, seed):
        self._h = lib().abo_rng_new(seed)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().abo_rng_free(self._h)
            self._h = None

    def u32(self):
        return lib().abo_rng_u32(self._h)

    def random_sample(self):
        return lib().abo_rng_double(self._h)

    def standard_normal(self):
        return lib().abo_rng_gauss(self._h)

    def standard_exponential(self):
        return lib().abo_rng_std_exponential(self._h)

    def randint(self, low, high):
        return lib().abo_rng_randint(self._h, low, high)


class OracleBook:
    """util/OrderBook.py restated (oracle/abides_oracle.c); notifications come back as int64 rows
    (t, recipient, kind, order_id, is_buy, qty, limit_price, fill_price, 0...)."""

    def __init__(self, stream_history=10):
        self._h = lib().abo_book_new(stream_history)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().abo_book_free(self._h)
            self._h = None

    def set_time(self, t):
        lib().abo_book_set_time(self._h, int(t))

    def limit(self, agent, order_id, is_buy, price, qty):
        lib().abo_book_limit(self._h, agent, order_id, int(is_buy), price, qty)

    def cancel(self, agent, order_id, is_buy, price):
        lib().abo_book_cancel(self._h, agent, order_id, int(is_buy), price)

    def modify(self, agent, order_id, is_buy, price, new_price, new_qty, new_order_id=None):
        lib().abo_book_modify(self._h, agent, order_id, int(is_buy), price,
                              order_id if new_order_id is None else new_order_id, new_price, new_qty)

    def inside(self, is_bid, depth):
        depth = min(depth, max(1, lib().abo_book_n_levels(self._h, int(is_bid))))
        buf = (C.c_int64 * (2 * depth))()
        n = lib().abo_book_inside(self._h, int(is_bid), depth, buf)
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]

    def level_orders(self, is_bid, level, max_orders=4096):
        buf = (C.c_int64 * (3 * max_orders))()
        n = lib().abo_book_level_orders(self._h, int(is_bid), level, buf, max_orders)
        return [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]

    @property
    def last_trade(self):
        v = lib().abo_book_last_trade(self._h)
        return None if v < 0 else v

    def transacted_volume(self, lookback_ns):
        return lib().abo_book_transacted_volume(self._h, lookback_ns)

    def n_levels(self, is_bid):
        return lib().abo_book_n_levels(self._h, int(is_bid))

    def n_resting(self):
        return lib().abo_book_n_resting(self._h)

    def take_notes(self):
        n = lib().abo_book_n_notes(self._h)
        arr = np.ctypeslib.as_array(lib().abo_book_notes(self._h), shape=(n * 13,)).reshape(n, 13).copy() if n else \
            np.zeros((0, 13), np.int64)
        lib().abo_book_clear_notes(self._h)
        return arr


class OracleSim:
    """config/sparse_zi_100.py / sparse_zi_1000.py + Kernel.runner restated (oracle/abides_oracle.c)."""

    def __init__(self, variant, seed, trace=0, pov_exec=None):
        """variant 100 / 1000: config/sparse_zi_*.py; variant 3: config/rmsc03.py.  pov_exec (variant 3 only): dict(pov, quantity, is_buy,
        start_ns, end_ns, freq_ns, lookback_ns) appends one POVExecutionAgent (agent/execution/baselines/pov_agent.py) as the last agent."""
        self._keep = None
        if variant == "config":                       # OracleSim.from_config
            return
        if variant == 3 and pov_exec:
            p = pov_exec
            self._h = lib().abo_sim_new_rmsc03_pov(seed, trace, float(p["pov"]), int(p["quantity"]), int(bool(p["is_buy"])), int(p["start_ns"]), int(p["end_ns"]),
                                                   int(p["freq_ns"]), int(p["lookback_ns"]))
        else:
            self._h = lib().abo_sim_new_rmsc03(seed, trace) if variant == 3 else lib().abo_sim_new_sparse_zi(variant, seed, trace)
        if not self._h:
            raise ValueError("unknown sparse_zi variant %r" % (variant,))
        self.variant, self.seed = variant, seed

    @classmethod
    def from_config(cls, cfg, seed, trace=0):
        """A simulation built from an abx_sim_config -- the ctypes struct the product takes (marl_optimal_execution_b200._lib.SimConfig, or any
        ctypes structure of that layout): same seed cascade as the config scripts, parameters and agent counts from `cfg`."""
        self = cls("config", seed)
        self._h = lib().abo_sim_new_config(C.addressof(cfg), int(seed), int(trace))
        if not self._h:
            raise ValueError("abo_sim_new_config rejected the configuration")
        self.variant = 3 if cfg.population == 1 else 1 if cfg.population == 3 else (100 if cfg.latency_model == 1 else 1000)
        self.seed = seed
        return self

    def set_external(self, bits, kinds, off, theta=None, lat_to=None, lat_from=None, sizes=None, wakes=None):
        """Replace every RandomState by a supplied list of standard variates (product stream order: symbol, kernel, latency model, global,
        agents 1..n-1) and install the start-of-run state drawn with them.  Used to re-run a Philox-seeded GPU simulation draw for draw."""
        keep = [np.ascontiguousarray(bits, np.uint64), np.ascontiguousarray(kinds, np.uint8), np.ascontiguousarray(off, np.int64)]
        opt = []
        for a, dt in ((theta, np.int32), (lat_to, np.float64), (lat_from, np.float64), (sizes, np.int32), (wakes, np.int64)):
            opt.append(None if a is None else np.ascontiguousarray(a, dt))
        keep += opt
        self._keep = keep                              # the C side keeps pointers into these arrays

        def ptr(a, ct):
            return None if a is None else a.ctypes.data_as(C.POINTER(ct))
        rc = lib().abo_sim_set_external(self._h, ptr(keep[0], C.c_uint64), ptr(keep[1], C.c_uint8), ptr(keep[2], C.c_int64), ptr(opt[0], C.c_int32),
                                        ptr(opt[1], C.c_double), ptr(opt[2], C.c_double), ptr(opt[3], C.c_int32), ptr(opt[4], C.c_int64))
        if rc != 0:
            raise ValueError("abo_sim_set_external failed")

    def rng_error(self):
        return lib().abo_sim_rng_error(self._h)

    def external_unread(self):
        return lib().abo_sim_external_unread(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().abo_sim_free(self._h)
            self._h = None

    def run(self):
        return lib().abo_sim_run(self._h)

    def pov_exec(self):
        out = np.zeros(3, dtype=np.int64)
        lib().abo_sim_pov_exec(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def start(self):
        lib().abo_sim_start(self._h)

    def stop(self):
        lib().abo_sim_stop(self._h)

    def run_until(self, until_ns):
        done = C.c_int(0)
        n = lib().abo_sim_run_until(self._h, int(until_ns), C.byref(done))
        return n, bool(done.value)

    @property
    def n_agents(self):
        return lib().abo_sim_n_agents(self._h)

    @property
    def n_pops(self):
        return lib().abo_sim_n_pops(self._h)

    def holdings(self):
        n = self.n_agents - 1
        out = np.zeros((n, 5), np.int64)
        lib().abo_sim_holdings(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def pop_hash(self):
        return lib().abo_sim_pop_hash(self._h)

    def hash_ckpt(self):
        n = lib().abo_sim_n_hash_ckpt(self._h)
        return np.ctypeslib.as_array(lib().abo_sim_hash_ckpt(self._h), shape=(n,)).copy() if n else np.zeros(0, np.uint64)

    def note_hash(self):
        return lib().abo_sim_note_hash(self._h)

    def snap_hash(self):
        return lib().abo_sim_snap_hash(self._h)

    def trace(self, which):
        w = {"pops": (0, 5), "ops": (1, 9), "notes": (2, 13), "snaps": (3, 16)}[which]
        p = C.POINTER(C.c_int64)()
        n = lib().abo_sim_trace(self._h, w[0], C.byref(p))
        if n == 0:
            return np.zeros((0, w[1]), np.int64)
        return np.ctypeslib.as_array(p, shape=(n * w[1],)).reshape(n, w[1]).copy()

    @property
    def n_streams(self):
        return lib().abo_sim_n_streams(self._h)

    def tape(self, stream):
        k = C.POINTER(C.c_uint8)()
        b = C.POINTER(C.c_uint64)()
        n = lib().abo_sim_tape(self._h, stream, C.byref(k), C.byref(b))
        if n == 0:
            return np.zeros(0, np.uint8), np.zeros(0, np.uint64)
        return (np.ctypeslib.as_array(k, shape=(n,)).copy(), np.ctypeslib.as_array(b, shape=(n,)).copy())

    def stream_seed(self, stream):
        return lib().abo_sim_stream_seed(self._h, stream)

    def global_tape(self):
        k = C.POINTER(C.c_uint8)()
        b = C.POINTER(C.c_uint64)()
        n = lib().abo_sim_global_tape(self._h, C.byref(k), C.byref(b))
        if n == 0:
            return np.zeros(0, np.uint8), np.zeros(0, np.uint64)
        return (np.ctypeslib.as_array(k, shape=(n,)).copy(), np.ctypeslib.as_array(b, shape=(n,)).copy())

    def agent_info(self, agent):
        out = np.zeros(4, np.int64)
        lib().abo_sim_agent_info(self._h, agent, out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def global_exp_tape(self):
        p = C.POINTER(C.c_double)()
        n = lib().abo_sim_global_exp_tape(self._h, C.byref(p))
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0)

    def theta(self, agent):
        out = np.zeros(20, np.int32)
        lib().abo_sim_theta(self._h, agent, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def latency_vectors(self):
        n = self.n_agents
        a, b = np.zeros(n), np.zeros(n)
        lib().abo_sim_latency_vectors(self._h, a.ctypes.data_as(C.POINTER(C.c_double)),
                                      b.ctypes.data_as(C.POINTER(C.c_double)))
        return a, b

    def zi_params(self, agent):
        out = np.zeros(4)
        lib().abo_sim_zi_params(self._h, agent, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def counter(self, which):
        names = ["limit", "cancel", "fills", "spread_queries", "max_queue", "max_bid_levels", "max_ask_levels",
                 "max_resting", "orders_allocated", "uniq"]
        return lib().abo_sim_counter(self._h, names.index(which))

    def book_l1(self):
        out = np.zeros(5, np.int64)
        lib().abo_sim_book_l1(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def fundamental(self):
        return lib().abo_sim_fundamental(self._h)


class OracleEnv:
    """ABIDESEnv (ABIDESEnv.py) restated: Exchange + MarketReplayAgent + DummyRLExecutionAgent under GymKernel.
    `stream` is an int64 [n,5] array of (t_ns, ORDER_ID, PRICE, SIZE, is_buy) rows."""

    def __init__(self, stream, quantity=1e5, order_level=2, trace=0, stop_ns=(16 * 3600 + 600) * 10 ** 9):
        """order_level 0 = no RL agent: config/marketreplay.py (stop_ns 16:01); one step() then runs the whole day."""
        self._stream = np.ascontiguousarray(stream, dtype=np.int64)
        self._h = lib().abo_env_new2(self._stream.ctypes.data_as(C.POINTER(C.c_int64)), len(self._stream), float(quantity),
                                     int(order_level), int(trace), int(stop_ns))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().abo_env_free(self._h)
            self._h = None

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        obs = np.zeros(9)
        done = C.c_int(0)
        n = lib().abo_env_step(self._h, a.ctypes.data_as(C.POINTER(C.c_double)), obs.ctypes.data_as(C.POINTER(C.c_double)),
                               C.byref(done))
        return obs[:n], None, int(done.value), None

    @property
    def n_pops(self):
        return lib().abo_env_n_pops(self._h)

    def pop_hash(self):
        return lib().abo_env_pop_hash(self._h)

    def note_hash(self):
        return lib().abo_env_note_hash(self._h)

    def snap_hash(self):
        return lib().abo_env_snap_hash(self._h)

    def hash_ckpt(self):
        n = lib().abo_env_n_hash_ckpt(self._h)
        return np.ctypeslib.as_array(lib().abo_env_hash_ckpt(self._h), shape=(n,)).copy() if n else np.zeros(0, np.uint64)

    def trace(self, which):
        w = {"pops": (0, 5), "ops": (1, 9), "notes": (2, 13), "snaps": (3, 16)}[which]
        p = C.POINTER(C.c_int64)()
        n = lib().abo_env_trace(self._h, w[0], C.byref(p))
        if n == 0:
            return np.zeros((0, w[1]), np.int64)
        return np.ctypeslib.as_array(p, shape=(n * w[1],)).reshape(n, w[1]).copy()

    def final(self):
        out = np.zeros(8)
        lib().abo_env_final(self._h, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def counter(self, which):
        names = ["max_queue", "max_bid_levels", "max_ask_levels", "max_resting", "uniq", "next_order_id"]
        return lib().abo_env_counter(self._h, names.index(which))


class OracleDDQNEnv(OracleEnv):
    """config/execution/marketreplay/execution_marketreplay_ddqn.py (-a rl) restated: Exchange + MarketReplayAgent + MomentumAgents +
    TWAPExecutionAgent + DDQLearningExecutionAgent under Kernel.runner, paused at every choose_action of the DDQN agent."""

    def __init__(self, stream, mom_sizes, n_twap=1, has_ddqn=True, is_buy=True, quantity=500000, h0_ns=10 * 3600 * 10 ** 9,
                 h_step_ns=30 * 10 ** 9, n_h=661, mom_wake_ns=20 * 10 ** 9, trace=0):
        self._stream = np.ascontiguousarray(stream, dtype=np.int64)
        ms = np.ascontiguousarray(mom_sizes, dtype=np.int64)
        self.n_agents = 2 + len(ms) + n_twap + (1 if has_ddqn else 0)
        self.n_exec = n_twap + (1 if has_ddqn else 0)
        self._h = lib().abo_dq_new(self._stream.ctypes.data_as(C.POINTER(C.c_int64)), len(self._stream), len(ms), ms.ctypes.data_as(C.POINTER(C.c_int64)),
                                   int(n_twap), int(bool(has_ddqn)), int(bool(is_buy)), int(quantity), int(h0_ns), int(h_step_ns), int(n_h), int(mom_wake_ns), int(trace))
        if not self._h:
            raise ValueError("abo_dq_new rejected the configuration")

    def set_schedule(self, k, qty):
        """Baseline execution agent k places qty[bin] instead of the TWAP child quantity (VWAPExecutionAgent.generate_schedule, vwap_agent.py:48-62)."""
        q = np.ascontiguousarray(qty, dtype=np.int64)
        if lib().abo_dq_set_schedule(self._h, int(k), q.ctypes.data_as(C.POINTER(C.c_int64)), len(q)) != 0:
            raise ValueError("abo_dq_set_schedule rejected the schedule")

    def step(self, action):
        """-> (obs8 = 6 features + 2 digitised state entries, trans6 = finalised (s, a, s', r) of the previous tick, reward, done)"""
        out8, tr = np.zeros(8), np.zeros(6)
        rew, done = C.c_double(0), C.c_int(0)
        lib().abo_dq_step(self._h, int(action), out8.ctypes.data_as(C.POINTER(C.c_double)), tr.ctypes.data_as(C.POINTER(C.c_double)), C.byref(rew), C.byref(done))
        return out8, tr, float(rew.value), int(done.value)

    def error(self):
        return lib().abo_dq_error(self._h)

    def series(self, which):
        w = {"price_path": (0, 1), "experience": (1, 6), "step_reward_hist": (2, 1), "action_hist": (3, 1)}[which]
        p = C.POINTER(C.c_double)()
        n = lib().abo_dq_series(self._h, w[0], C.byref(p))
        if n == 0:
            return np.zeros((0, w[1])) if w[1] > 1 else np.zeros(0)
        a = np.ctypeslib.as_array(p, shape=(n * w[1],)).copy()
        return a.reshape(n, w[1]) if w[1] > 1 else a

    def holdings(self):
        out = np.zeros((self.n_agents - 1, 5), dtype=np.int64)
        lib().abo_dq_holdings(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)))
        return out

    def exec_final(self, k):
        out = np.zeros(5)
        lib().abo_dq_exec_final(self._h, int(k), out.ctypes.data_as(C.POINTER(C.c_double)))
        return out
