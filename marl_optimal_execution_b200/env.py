"""ABIDESEnv: batched, GPU-resident mirror of the reference's gym surface (ABIDESEnv.py:7-57).

    env = ABIDESEnv(stream, n_envs=9472)      # whole waves: a multiple of 148 SMs x 16 resident environments; stream: int64 [n,5] rows (t_ns, ORDER_ID, PRICE cents, SIZE, is_buy)
    env.reset()
    obs, reward, done, info = env.step(actions)   # actions [n_envs, 3] in [0,1]: (x_hat, o_hat_1, o_hat_2)

Same names and tuple order as the reference; everything is batched over independent environments that all replay the
same LOBSTER order stream (Exchange + MarketReplayAgent + DummyRLExecutionAgent under GymKernel).  `actions` may be a
CUDA torch tensor (fp64; obs/reward/done come back as CUDA tensors, nothing touches the host) or a numpy array (host
buffers through abx_env_step_host).  reward is 0 where the reference returns None (its get_reward ends in `return None`,
agent/execution/rl/dummy_rl_execution_agent.py:325-352); obs rows are 0 where the reference returns [].
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DqConfig, EnvConfig


def env_config(lib=None, **overrides):
    L = lib or _lib.load()
    cfg = EnvConfig()
    _lib.check(L, L.abx_env_config_default(C.byref(cfg)), "abx_env_config_default")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError("abx_env_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg


def _pack_days(stream):
    """One stream (int64 [n, 5]) or a list of them (one per replayed day; environment e replays day e % n_days) -> (rows, offsets)."""
    days = [stream] if isinstance(stream, np.ndarray) else list(stream)
    arrs = [np.ascontiguousarray(d, dtype=np.int64) for d in days]
    for a in arrs:
        if a.ndim != 2 or a.shape[1] != 5:
            raise ValueError("stream must be int64 [n, 5]: (t_ns, ORDER_ID, PRICE, SIZE, is_buy)")
    off = np.zeros(len(arrs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(a) for a in arrs])
    return np.ascontiguousarray(np.concatenate(arrs, axis=0)), off


def load_lobster_fixture(path):
    """tests/golden/*.npz written by tools/record_reference_env.py: the stream as the reference parsed it."""
    return np.ascontiguousarray(np.load(path)["stream"], dtype=np.int64)


def load_lobster_csv(path, start_s=34200, end_s=57600):
    """A LOBSTER message file -> the replayed stream int64 [n, 5] = (t_ns since midnight, ORDER_ID, PRICE cents, SIZE, is_buy), row for row
    what LOBSTEROrdersProcessor.processOrders builds (agent/examples/MarketReplayAgent.py:196-216):
      columns TIMESTAMP(s), EVENT_TYPE, ORDER_ID, SIZE, PRICE(x1e4), BUY_SELL_FLAG(+1/-1), every event type kept (:198-202);
      TIMESTAMP = start_time + pd.to_timedelta(seconds, "s") - 09:30 (:206-208) -- pandas splits the float into whole seconds and a
        fraction rounded to 9 decimals, and truncates fraction * 1e9 to integer nanoseconds;
      PRICE = int(float(price) / 100) (:210-211: truncation to cents); SIZE int; rows with mkt_open <= t < mkt_close (:212);
      grouped per timestamp in ascending order, file order inside a group (:216)."""
    raw = np.loadtxt(path, delimiter=",", dtype=np.float64, usecols=(0, 2, 3, 4, 5), ndmin=2)
    t = raw[:, 0]
    base = np.trunc(t)
    t_ns = base.astype(np.int64) * 10 ** 9 + (np.round(t - base, 9) * 1e9).astype(np.int64)
    rows = np.stack([t_ns, raw[:, 1].astype(np.int64), (raw[:, 3] / 100).astype(np.int64), raw[:, 2].astype(np.int64),
                     (raw[:, 4].astype(np.int64) == 1).astype(np.int64)], axis=1)
    rows = rows[(t_ns >= int(start_s) * 10 ** 9) & (t_ns < int(end_s) * 10 ** 9)]
    return np.ascontiguousarray(rows[np.argsort(rows[:, 0], kind="stable")])


def lobster_message_path(ticker, date, data_root="data/lobster", level=1, dated_folder=False):
    """Where the reference looks for a day's message file: agent_config.py:63-64 (ABIDESEnv: LOBSTER_SampleFile_{ticker}_{level}/) or, with
    dated_folder=True, config/marketreplay.py:91-92 (LOBSTER_SampleFile_{ticker}_{date}_{level}/)."""
    import os
    folder = "LOBSTER_SampleFile_{}_{}_{}".format(ticker, date, level) if dated_folder else "LOBSTER_SampleFile_{}_{}".format(ticker, level)
    return os.path.join(data_root, folder, "{}_{}_34200000_57600000_message_{}.csv".format(ticker, date, level))


def _load_days(ticker, dates, data_root, level):
    days = []
    for d in dates:
        path = lobster_message_path(ticker, d, data_root, level)
        rows = load_lobster_csv(path)
        if len(rows) == 0:          # the reference's constructor fails the same way: wakeup_times[0] of an empty dict (MarketReplayAgent.py:177)
            raise IndexError("no orders between 09:30 and 16:00 in %s" % path)
        days.append(rows)
    return days


class Box:
    """The two attributes of gym.spaces.Box the reference's callers read (ABIDESEnv.py:22-25): low / high (+ shape, dtype, sample, contains)."""

    def __init__(self, low, high):
        self.low = np.asarray(low, dtype=np.float32)
        self.high = np.asarray(high, dtype=np.float32)
        self.shape, self.dtype = self.low.shape, self.low.dtype

    def sample(self, rng=None):
        return (rng or np.random).uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class ABIDESEnv:
    OBS_SIZE = 9          # get_observation returns 9 values (get_observation_space_size says 10)

    def __init__(self, ticker, date=None, log_dir=None, seed=None, n_envs=1, device=0, cfg=None, lib_path=None, data_root="data/lobster", level=1):
        """Positionally the reference's constructor: ABIDESEnv(ticker, date, log_dir=None, seed=None) (ABIDESEnv.py:8) replays the day's LOBSTER message
        file found under `data_root` where agent_config.py:63-64 looks for it (`date` may be a list of days).  log_dir and seed are accepted and unused
        on this path (no logging, no random draws: zero latency, noise [1.0]).  Batched form: the first argument may instead be the parsed stream
        itself -- an int64 [n, 5] array of (t_ns, ORDER_ID, PRICE cents, SIZE, is_buy) rows, or a list of such arrays, one per replayed day."""
        if isinstance(ticker, str):
            if date is None:
                raise TypeError("ABIDESEnv(ticker, date, ...): date is required")
            self.ticker, self.date, self.log_dir, self.seed = ticker, date, log_dir, seed
            stream = _load_days(ticker, [date] if isinstance(date, str) else list(date), data_root, level)
        else:
            stream = ticker
        self._L = _lib.load(lib_path)
        self.cfg = cfg or env_config(self._L)
        self.n_envs = int(n_envs)
        self.device = int(device)
        self.action_size = int(self.cfg.order_level) + 1                  # get_action_space_size :128-134
        # ABIDESEnv.py:18-25: actions in [0, 1]^(k+1); the observation Box is built from get_observation_space_size() = 10 zeros for low AND
        # high (dummy_rl_execution_agent.py:317-322) although get_observation returns 9 values -- mirrored as is
        self.action_space = Box([0.0] * self.action_size, [1.0] * self.action_size)
        self.observation_space = Box([0] * 10, [0] * 10)
        st, off = _pack_days(stream)
        self.n_days = len(off) - 1
        self._h = C.c_void_p()
        _lib.check(self._L, self._L.abx_env_create_days(C.byref(self.cfg), st.ctypes.data_as(C.POINTER(C.c_int64)), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                                        self.n_days, self.n_envs, self.device, C.byref(self._h)), "abx_env_create_days")
        self._torch_out = None

    @classmethod
    def from_lobster(cls, ticker, date, n_envs=1, data_root="data/lobster", level=1, **kw):
        """The reference's constructor arguments (ABIDESEnv(ticker, date), ABIDESEnv.py:8-17): replays the day's LOBSTER message file found
        where agent_config.py:63-64 looks for it under `data_root`.  `date` may be a list of 'yyyy-mm-dd' strings (environment e replays
        day e % n_days).  log_dir / seed of the reference have no effect on this path (no random draws, no logging)."""
        return cls(ticker, date, n_envs=n_envs, data_root=data_root, level=level, **kw)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.abx_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, mask=None, advance_day=False, stream=None):
        """ABIDESEnv.reset() (ABIDESEnv.py:51-57).  Like the reference it returns None: the first observation comes from the first step().
        mask (bool / uint8 [n_envs], numpy or CUDA tensor): only those environments start over -- the reference's reset is per environment object;
        advance_day moves each of them on to its next replayed day (environment e replays day (e + resets) % n_days)."""
        if mask is None:
            _lib.check(self._L, self._L.abx_env_reset(self._h, stream), "abx_env_reset")
            return None
        _lib.check(self._L, self._L.abx_env_reset_mask(self._h, self._mask_ptr(mask), int(bool(advance_day)), stream), "abx_env_reset_mask")
        return None

    def _mask_ptr(self, mask):
        """Device pointer of a uint8 [n_envs] mask (a CUDA tensor is used in place; a host array is copied through torch; the CPU emulation of the test-suite takes the host pointer)."""
        try:
            import torch
        except ImportError:      # pragma: no cover
            torch = None
        if torch is not None and isinstance(mask, torch.Tensor) and mask.is_cuda:
            self._mask_keep = mask.to(torch.uint8).contiguous().view(self.n_envs)
            return C.c_void_p(self._mask_keep.data_ptr())
        m = np.ascontiguousarray(np.asarray(mask).astype(np.uint8).reshape(self.n_envs))
        if torch is not None and torch.cuda.is_available() and self._L.abx_device_count() > 0:
            self._mask_keep = torch.from_numpy(m).to("cuda:%d" % self.device)
            return C.c_void_p(self._mask_keep.data_ptr())
        self._mask_keep = m
        return C.c_void_p(m.ctypes.data)

    def set_auto_reset(self, mode):
        """After every step, environments whose episode has ended (done == 1 in that step's output) are reset before the next step:
        0 / False off, 1 / True restart the same day, 2 / "next_day" move on to the next replayed day."""
        mode = 2 if mode == "next_day" else int(mode)
        _lib.check(self._L, self._L.abx_env_set_auto_reset(self._h, mode), "abx_env_set_auto_reset")

    reuse_outputs = False      # True: the CUDA path returns the same three tensors on every step (no allocation in a tight loop; bench.py sets it)

    def step(self, actions, stream=None, out=None):
        """(obs [n_envs, 9], reward [n_envs], done [n_envs], None), as ABIDESEnv.step (ABIDESEnv.py:30-49).  actions: [n_envs, order_level + 1]
        (x_hat, o_hat_1 .. o_hat_k) in [0, 1].  CUDA tensors in -> fresh CUDA tensors out (or `out=(obs, reward, done)` to fill preallocated ones,
        or set `reuse_outputs`)."""
        try:
            import torch
            is_torch = isinstance(actions, torch.Tensor)
        except ImportError:      # pragma: no cover
            is_torch = False
        if is_torch and actions.is_cuda:
            import torch
            a = actions.to(torch.float64).reshape(self.n_envs, -1)
            if a.shape[1] != self.action_size and a.shape[1] != 3:
                raise ValueError("actions must be [n_envs, %d] (order_level + 1)" % self.action_size)
            if a.shape[1] < 3:                                            # the device layout always has three columns (x_hat, o_hat_1, o_hat_2)
                a = torch.cat([a, torch.zeros(self.n_envs, 3 - a.shape[1], dtype=torch.float64, device=a.device)], dim=1)
            a = a.contiguous()
            if out is not None:
                obs, rew, done = out
            elif self.reuse_outputs:
                if self._torch_out is None:
                    self._torch_out = (torch.zeros(self.n_envs, 9, dtype=torch.float64, device=a.device),
                                       torch.zeros(self.n_envs, dtype=torch.float64, device=a.device),
                                       torch.zeros(self.n_envs, dtype=torch.uint8, device=a.device))
                obs, rew, done = self._torch_out
            else:                                                         # fresh tensors per step, like the reference's fresh arrays
                obs = torch.empty(self.n_envs, 9, dtype=torch.float64, device=a.device)
                rew = torch.empty(self.n_envs, dtype=torch.float64, device=a.device)
                done = torch.empty(self.n_envs, dtype=torch.uint8, device=a.device)
            sp = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream) if stream is None else stream
            self._keep_actions = a                                        # `a` may be a temporary and `stream` need not be torch's current stream: it stays referenced until the next step replaces it
            _lib.check(self._L, self._L.abx_env_step(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(obs.data_ptr()),
                                                     C.c_void_p(rew.data_ptr()), C.c_void_p(done.data_ptr()), sp), "abx_env_step")
            return obs, rew, done, None
        a = np.asarray(actions, dtype=np.float64).reshape(self.n_envs, -1)
        if a.shape[1] != self.action_size and a.shape[1] != 3:
            raise ValueError("actions must be [n_envs, %d] (order_level + 1)" % self.action_size)
        if a.shape[1] < 3:
            a = np.concatenate([a, np.zeros((self.n_envs, 3 - a.shape[1]))], axis=1)
        a = np.ascontiguousarray(a)
        obs = np.zeros((self.n_envs, 9))
        rew = np.zeros(self.n_envs)
        done = np.zeros(self.n_envs, dtype=np.uint8)
        _lib.check(self._L, self._L.abx_env_step_host(self._h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data,
                                                      stream), "abx_env_step_host")
        return obs, rew, done, None

    def step_host_buffers(self, actions_ptr, obs_ptr, reward_ptr, done_ptr, stream=None):
        """Raw-pointer form of the host-buffer step (pinned buffers make the copies asynchronous)."""
        _lib.check(self._L, self._L.abx_env_step_host(self._h, C.c_void_p(actions_ptr), C.c_void_p(obs_ptr), C.c_void_p(reward_ptr),
                                                      C.c_void_p(done_ptr), stream), "abx_env_step_host")

    # shared with BatchedSim: per-environment counters and traces
    def stats(self, stream=None):
        out = np.zeros(self.n_envs, dtype=_lib.STATS_DTYPE)
        _lib.check(self._L, self._L.abx_sim_stats(self._h, out.ctypes.data, stream), "abx_sim_stats")
        return out

    def trace(self, env, stream=None):
        cap = int(self.cfg.trace_cap)
        out = np.zeros(max(cap, 1), dtype=_lib.TRACE_DTYPE)
        n = C.c_int32(0)
        _lib.check(self._L, self._L.abx_sim_trace(self._h, int(env), out.ctypes.data, cap, C.byref(n), stream), "abx_sim_trace")
        return out[: n.value]

    def split_trace(self, env):
        from .sim import BatchedSim
        return BatchedSim.split_trace(self, env)

    @property
    def launch_count(self):
        return int(self._L.abx_sim_launch_count(self._h))


def dq_config(lib=None, **overrides):
    """abx_dq_config with the defaults of config/execution/marketreplay/execution_marketreplay_ddqn.py (BUY 500 000 from 10:00
    over 330 min at 30 s; 7 MomentumAgents, one TWAP agent, the DDQN agent)."""
    L = lib or _lib.load()
    cfg = DqConfig()
    _lib.check(L, L.abx_dq_config_default(C.byref(cfg)), "abx_dq_config_default")
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError("abx_dq_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg


def vwap_schedule(volume_profile, quantity):
    """VWAPExecutionAgent.generate_schedule (agent/execution/baselines/vwap_agent.py:48-62): per-bin child quantities round(profile[bin] * quantity) (Python's
    round: half to even) for a volume profile given as one fraction per horizon bin."""
    return np.array([int(round(float(f) * quantity)) for f in volume_profile], dtype=np.int32)


def synthetic_volume_profile(n_bins):
    """The U shape of VWAPExecutionAgent.synthetic_volume_profile (vwap_agent.py:64-78: x^2 + 2x + 2 over x = -n/2 .. n/2 - 1, normalised) over n_bins bins."""
    w = np.array([x * x + 2 * x + 2 for x in range(int(-n_bins / 2), int(-n_bins / 2) + n_bins)], dtype=np.float64)
    return w * (1.0 / w.sum())


class DDQNExecutionEnv(ABIDESEnv):
    """The reference's DDQN execution simulation (config/execution/marketreplay/execution_marketreplay_ddqn.py, -a rl) as a
    batched decision process: Exchange + MarketReplayAgent + MomentumAgents + TWAPExecutionAgent + DDQLearningExecutionAgent.

        env = DDQNExecutionEnv(stream, n_envs=9472); env.reset(seeds)      # whole waves of 148 x 16 resident environments
        obs, trans, reward, done = env.step(None)            # runs to the first decision tick (10:00)
        obs, trans, reward, done = env.step(actions)         # actions int32 [n_envs] in 0..23 (ACTIONS, ddqlearning_execution_agent.py:24-37)

    obs [n_envs, 8]: the 6 features of get_observation (:332) followed by the 2 digitised entries the Q-network sees (:334);
    trans [n_envs, 6]: the finalised experience entry of the previous tick (s0, s1, a, s'0, s'1, r) -- what the reference stores
    in self.experience (:251,540-541,572-573), r NaN where it holds None; reward: sum of the per-fill rewards (:411-447) since
    the previous decision; done: the event loop ended (the reference's kernel.runner returned)."""
    OBS_SIZE = 8
    N_ACTIONS = 24

    def __init__(self, ticker, date=None, log_dir=None, seed=None, n_envs=1, device=0, cfg=None, lib_path=None, data_root="data/lobster", level=1):
        """(ticker, date) as the reference's config takes them (-t, -d), or the parsed stream(s) as the first argument (see ABIDESEnv)."""
        if isinstance(ticker, str):
            if date is None:
                raise TypeError("DDQNExecutionEnv(ticker, date, ...): date is required")
            self.ticker, self.date, self.log_dir, self.seed = ticker, date, log_dir, seed
            stream = _load_days(ticker, [date] if isinstance(date, str) else list(date), data_root, level)
        else:
            stream = ticker
        self._L = _lib.load(lib_path)
        self.cfg = cfg or dq_config(self._L)
        self.n_envs, self.device = int(n_envs), int(device)
        self.n_agents = 2 + int(self.cfg.n_momentum) + int(self.cfg.n_twap) + (1 if self.cfg.has_ddqn else 0)
        self.n_exec = int(self.cfg.n_twap) + (1 if self.cfg.has_ddqn else 0)
        st, off = _pack_days(stream)
        self.n_days = len(off) - 1
        self._h = C.c_void_p()
        _lib.check(self._L, self._L.abx_dq_create_days(C.byref(self.cfg), st.ctypes.data_as(C.POINTER(C.c_int64)), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                                       self.n_days, self.n_envs, self.device, C.byref(self._h)), "abx_dq_create_days")
        self._torch_out = None

    @classmethod
    def from_lobster(cls, ticker, date, n_envs=1, data_root="data/lobster", level=1, **kw):
        """(ticker, date) as config/execution/marketreplay/execution_marketreplay_ddqn.py takes them (-t, -d): the day's LOBSTER message
        file(s) under `data_root`, parsed like LOBSTEROrdersProcessor (load_lobster_csv)."""
        return cls(ticker, date, n_envs=n_envs, data_root=data_root, level=level, **kw)

    def set_schedule(self, k, qty):
        """Baseline execution agent k (0 .. n_twap-1) trades a per-bin schedule instead of the TWAP quantity: the reference's VWAPExecutionAgent
        (agent/execution/baselines/vwap_agent.py:48-62) -- `qty[b] = round(volume_profile[bin b] * quantity)`, see `vwap_schedule`.  Call before reset()."""
        q = np.ascontiguousarray(qty, dtype=np.int32)
        _lib.check(self._L, self._L.abx_dq_set_schedule(self._h, int(k), q.ctypes.data_as(C.POINTER(C.c_int32)), len(q)), "abx_dq_set_schedule")

    def reset(self, seeds=None, mom_sizes=None, stream=None, mask=None, advance_day=False):
        """Whole-batch reset with per-environment seeds (or recorded MomentumAgent sizes); with `mask`, only those environments start over, keeping the
        seeds / sizes of the last whole-batch reset (seed + episode number for the Philox-drawn sizes), optionally on their next replayed day."""
        if mask is not None:
            _lib.check(self._L, self._L.abx_env_reset_mask(self._h, self._mask_ptr(mask), int(bool(advance_day)), stream), "abx_env_reset_mask")
            return None
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64).reshape(self.n_envs)
        ms = None if mom_sizes is None else np.ascontiguousarray(mom_sizes, dtype=np.int32).reshape(self.n_envs, int(self.cfg.n_momentum))
        _lib.check(self._L, self._L.abx_dq_reset(self._h, None if sd is None else sd.ctypes.data, None if ms is None else ms.ctypes.data, stream), "abx_dq_reset")
        self._keep = (sd, ms)
        return None

    def step(self, actions, stream=None, out=None):
        try:
            import torch
            is_torch = isinstance(actions, torch.Tensor)
        except ImportError:      # pragma: no cover
            is_torch = False
        if is_torch and actions.is_cuda:
            import torch
            a = actions.to(torch.int32).contiguous().view(self.n_envs)
            if out is not None:
                obs, trans, rew, done = out
            elif self.reuse_outputs:
                if self._torch_out is None:
                    self._torch_out = (torch.zeros(self.n_envs, 8, dtype=torch.float64, device=a.device), torch.zeros(self.n_envs, 6, dtype=torch.float64, device=a.device),
                                       torch.zeros(self.n_envs, dtype=torch.float64, device=a.device), torch.zeros(self.n_envs, dtype=torch.uint8, device=a.device))
                obs, trans, rew, done = self._torch_out
            else:
                obs, trans = torch.empty(self.n_envs, 8, dtype=torch.float64, device=a.device), torch.empty(self.n_envs, 6, dtype=torch.float64, device=a.device)
                rew, done = torch.empty(self.n_envs, dtype=torch.float64, device=a.device), torch.empty(self.n_envs, dtype=torch.uint8, device=a.device)
            sp = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream) if stream is None else stream
            _lib.check(self._L, self._L.abx_dq_step(self._h, C.c_void_p(a.data_ptr()), C.c_void_p(obs.data_ptr()), C.c_void_p(trans.data_ptr()),
                                                    C.c_void_p(rew.data_ptr()), C.c_void_p(done.data_ptr()), sp), "abx_dq_step")
            return obs, trans, rew, done
        a = None if actions is None else np.ascontiguousarray(np.asarray(actions, dtype=np.int32).reshape(self.n_envs))
        obs, trans = np.zeros((self.n_envs, 8)), np.zeros((self.n_envs, 6))
        rew, done = np.zeros(self.n_envs), np.zeros(self.n_envs, dtype=np.uint8)
        _lib.check(self._L, self._L.abx_dq_step_host(self._h, None if a is None else a.ctypes.data, obs.ctypes.data, trans.ctypes.data, rew.ctypes.data,
                                                     done.ctypes.data, stream), "abx_dq_step_host")
        return obs, trans, rew, done

    def holdings(self, env, stream=None):
        """(rows (agent id, shares, cash, last_trade, open orders | -1), rows per execution agent (remaining qty, arrival, fills, remaining_time, t))"""
        out = np.zeros((self.n_agents - 1, 5), dtype=np.int64)
        ex = np.zeros((max(self.n_exec, 1), 5))
        _lib.check(self._L, self._L.abx_dq_holdings(self._h, int(env), out.ctypes.data, ex.ctypes.data, stream), "abx_dq_holdings")
        return out, ex[: self.n_exec]
