"""Shared cases for per-environment reset / auto-reset / order_level 1: run on the CPU emulation (CPU suite) and on the CUDA library (GPU suite)."""
import os

import numpy as np

from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.env import ABIDESEnv, DDQNExecutionEnv, dq_config, env_config
from oracle.oracle import OracleDDQNEnv, OracleEnv

DAYS = ("env_IBM_2003-01-14_s789.npz", "env_IBM_2003-01-15_s4242.npz", "ddqn_IBM_2003-01-16_s99_sell.npz")


def _days(golden_dir):
    return [np.load(os.path.join(golden_dir, f))["stream"] for f in DAYS]


def abidesenv_masked_reset_and_rotation(golden_dir, lib_path=None, n_steps=40):
    """Four environments over three days.  After n_steps, environments 1 and 3 are reset: 1 restarts its day, 3 moves on to its next day (day 0 -> the
    environment index 3 replays day (3 + 1) % 3 = 1).  Every environment must keep matching an oracle of the day it is on, step for step."""
    days = _days(golden_dir)
    L = _lib.load(lib_path)
    env = ABIDESEnv(days, n_envs=4, cfg=env_config(L, hash_pops=1), lib_path=lib_path)
    env.reset()
    orc = [OracleEnv(days[e % 3]) for e in range(4)]
    rs = np.random.RandomState(3)

    def step_all():
        a = np.stack([np.array([rs.uniform(0, 0.04), rs.uniform(), rs.uniform()]) for _ in range(4)])
        obs, _, done, _ = env.step(a)
        for e, o in enumerate(orc):
            x, _, od, _ = o.step(a[e])
            ref = np.zeros(9); ref[: len(x)] = x
            assert np.allclose(obs[e], ref, rtol=1e-9, atol=1e-12) and int(done[e]) == od, e
    for _ in range(n_steps):
        step_all()
    env.reset(mask=np.array([0, 1, 0, 0], dtype=np.uint8))                       # same day again
    env.reset(mask=np.array([0, 0, 0, 1], dtype=np.uint8), advance_day=True)     # on to the next day
    orc[1] = OracleEnv(days[1]); orc[3] = OracleEnv(days[(3 + 1) % 3])
    for _ in range(n_steps):
        step_all()
    st = env.stats()
    assert [int(h) for h in st["pop_hash"]] == [o.pop_hash() for o in orc] and (st["flags"] & _lib.F_ERROR_MASK == 0).all()
    assert [int(m) for m in st["messages"]] == [o.n_pops for o in orc]
    env.close()


def ddqn_auto_reset(golden_dir, lib_path=None, ticks_after=12):
    """DDQN execution shape with auto-reset "next day": an episode is run to its end (660 ticks + the closing events), the finished environments restart
    on their next day before the following step, and from there on match a fresh oracle of that day."""
    gs = [np.load(os.path.join(golden_dir, f)) for f in DAYS]
    days = [g["stream"] for g in gs]
    ms = gs[2]["mom_sizes"].astype(np.int32)
    L = _lib.load(lib_path)
    env = DDQNExecutionEnv(days, n_envs=3, cfg=dq_config(L, hash_pops=1), lib_path=lib_path)
    env.set_auto_reset("next_day")
    env.reset(mom_sizes=np.tile(ms, (3, 1)))
    rs = np.random.RandomState(8)
    obs, trans, rew, done = env.step(None)
    n = 0
    while not done.all():
        a = int(rs.randint(0, 24))
        obs, trans, rew, done = env.step(np.full(3, a, dtype=np.int32))
        n += 1
        assert n < 700
    assert n == 660
    orc = [OracleDDQNEnv(days[(e + 1) % 3], ms) for e in range(3)]                # every environment is now on its next day
    obs, trans, rew, done = env.step(None)                                        # first step of the new episodes: up to the first decision tick
    outs = [o.step(0) for o in orc]
    for k in range(ticks_after):
        for e in range(3):
            assert np.allclose(obs[e], outs[e][0], rtol=1e-9, atol=1e-12), (k, e)
        a = int(rs.randint(0, 24))
        obs, trans, rew, done = env.step(np.full(3, a, dtype=np.int32))
        outs = [o.step(a) for o in orc]
    st = env.stats()
    assert [int(h) for h in st["pop_hash"]] == [o.pop_hash() for o in orc] and (st["flags"] & _lib.F_ERROR_MASK == 0).all() and not done.any()
    env.close()


def order_level_one(golden_dir, lib_path=None, n_steps=30):
    """order_level = 1: the action has two entries (x_hat, o_hat_1); one limit order per step at the best level."""
    g = np.load(os.path.join(golden_dir, DAYS[0]))
    L = _lib.load(lib_path)
    env = ABIDESEnv(g["stream"], n_envs=2, cfg=env_config(L, order_level=1, hash_pops=1), lib_path=lib_path)
    assert env.action_space.shape == (2,)
    env.reset()
    o = OracleEnv(g["stream"], order_level=1)
    rs = np.random.RandomState(1)
    for _ in range(n_steps):
        a = np.array([rs.uniform(0, 0.05), rs.uniform()])
        obs, _, done, _ = env.step(np.tile(a, (2, 1)))
        x, _, od, _ = o.step(a)
        ref = np.zeros(9); ref[: len(x)] = x
        assert np.allclose(obs[0], ref, rtol=1e-9, atol=1e-12) and np.array_equal(obs[0], obs[1])
    st = env.stats()
    assert (st["pop_hash"] == np.uint64(o.pop_hash())).all() and (st["flags"] & _lib.F_ERROR_MASK == 0).all()
    env.close()


def generated_ids_skip_explicit_ids(lib_path=None):
    """util/order/Order.py:35-42: generateOrderId hands out the smallest id no Order has used yet -- explicit ORDER_IDs of the replayed stream included,
    from the row that first carries them.  A synthetic day whose explicit ids (3, 5, 6, 9) lie inside the range the RL agent's generated ids run through:
    every exchange message (order ids included) must equal the oracle's, which keeps the reference's used-id list.  An explicit id that the generator
    reaches BEFORE its first row (the reference would then run two orders under one id) raises ABX_F_ID_RANGE."""
    NS, T = 10 ** 9, 34200 * 10 ** 9
    rows = [(T + 1 * NS, 3, 9990, 500, 1), (T + 2 * NS, 5, 10010, 500, 0), (T + 3 * NS, 6, 9980, 300, 1), (T + 4 * NS, 9, 10020, 300, 0)]
    rows += [(T + (700 + 40 * k) * NS, 3 if k % 2 else 5, 9990 if k % 2 else 10010, 500 - k, k % 2) for k in range(1, 30)]       # modifies of ids 3 / 5 through the day
    stream = np.array(rows, dtype=np.int64)
    L = _lib.load(lib_path)
    env = ABIDESEnv(stream, n_envs=2, cfg=env_config(L, trace_cap=20000, hash_pops=1), lib_path=lib_path)
    env.reset()
    o = OracleEnv(stream, trace=31)
    rs = np.random.RandomState(4)
    for _ in range(12):
        a = np.array([rs.uniform(0.001, 0.01), rs.uniform(), rs.uniform()])
        env.step(np.tile(a, (2, 1)))
        o.step(a)
    st = env.stats()
    p, nt, sn = env.split_trace(1)
    assert np.array_equal(p, o.trace("pops")) and np.array_equal(nt, o.trace("notes")) and np.array_equal(sn, o.trace("snaps"))
    ids = sorted(set(int(x) for x in nt[nt[:, 2] == 7][:, 3]))                  # ORDER_ACCEPTED ids: the stream's 3, 5, 6, 9 and generated ones around them
    assert {3, 5, 6, 9} <= set(ids) and {0, 1, 2, 4, 7, 8, 10} <= set(ids) and (st["flags"] & _lib.F_ID_RANGE == 0).all()
    env.close()
    late = np.array(rows + [(T + 20000 * NS, 12, 9970, 100, 1)], dtype=np.int64)                                          # id 12 first appears at 15:03: the generator gets there first
    env = ABIDESEnv(late, n_envs=1, cfg=env_config(L), lib_path=lib_path)
    env.reset()
    for _ in range(12):
        env.step(np.array([[0.005, 0.5, 0.5]]))
    assert (env.stats()["flags"] & _lib.F_ID_RANGE != 0).all()
    env.close()
