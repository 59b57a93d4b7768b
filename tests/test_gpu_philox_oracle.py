"""The BENCHMARKED kernel instantiations against the oracle (run on the B200 box, through the C ABI).

1. Philox mode (what bench.py times): every standard variate an environment draws is logged (cfg.draw_log_cap), the oracle -- built from the
   SAME abx_sim_config -- is re-run on exactly those draws (external tapes), and message count, pop hash, counters, L1, fundamental and the
   holdings of every agent must be equal.  The un-instrumented production kernel (abx_run_kernel<PHILOX, ., INSTR=false>) is then run on the
   same seeds and must reproduce every counter and holding of the instrumented run: production kernel == oracle, bit for bit.
2. Tape mode without instrumentation (INSTR=false) against the oracle.
3. Tape parity at PRODUCTION OCCUPANCY: one full wave of 148 SMs x 16 resident one-warp CTAs (2 368 environments) replaying 8 recorded runs
   round robin; every environment's pop hash, counters and holdings equal its oracle's -- for the default library and for the
   -DABX_STRICT_SYNC build (every on-chip sync a real __syncwarp()).  ZI, rmsc03, ABIDESEnv and DDQN shapes.
"""
import os

import numpy as np
import pytest
import torch

from helpers import assert_env_equals_oracle, oracle_rerun_of_philox_env, oracle_tapes
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.build import LIB_STRICT
from marl_optimal_execution_b200.sim import BatchedSim, rmsc01_config, rmsc02_config, rmsc03_config, sparse_zi_config
from oracle.oracle import OracleSim, TRACE_ALL

pytestmark = pytest.mark.gpu

WAVE = 148 * 16                       # one-warp CTAs resident on one B200 at the kernels' launch bounds
STAT_FIELDS = ["messages", "now_ns", "limit_orders", "cancels", "fills", "spread_queries", "max_queue", "n_bid_levels", "n_ask_levels", "n_resting",
               "best_bid", "best_bid_qty", "best_ask", "best_ask_qty", "last_trade", "fundamental", "flags", "uniq", "orders_allocated", "sum_shares", "sum_cash"]
LIBS = [pytest.param(None, id="default"), pytest.param(LIB_STRICT, id="strict_sync")]


def _philox_vs_oracle(make_cfg, n_envs, seed0, holdings_cols, skip_flag=0):
    seeds = np.arange(n_envs, dtype=np.uint64) + np.uint64(seed0)
    sim = BatchedSim(make_cfg(hash_pops=1), n_envs)
    sim.reset(seeds)
    init = [sim.agent_init(e) for e in range(n_envs)]
    sim.run()
    sim.finalize()
    st_i = sim.stats()
    assert (st_i["flags"] & _lib.F_TRACE_OVERFLOW == 0).all()
    hold_i, compared = [], 0
    for e in range(n_envs):
        hold_i.append(sim.holdings(e))
        if int(st_i["flags"][e]) & skip_flag:
            continue
        o, n = oracle_rerun_of_philox_env(sim, e, init[e], 0)
        assert_env_equals_oracle(sim, e, o, n, st_i, traces=False, holdings_cols=holdings_cols)
        compared += 1
    sim.close()
    assert compared >= n_envs * 3 // 4
    # the production instantiation: same seeds, no instrumentation compiled in
    sim = BatchedSim(make_cfg(hash_pops=0, draw_log_cap=0), n_envs)
    sim.reset(seeds)
    sim.run()
    sim.finalize()
    st_p = sim.stats()
    for f in STAT_FIELDS:
        assert np.array_equal(st_p[f], st_i[f]), f
    assert (st_p["pop_hash"] != st_i["pop_hash"]).all()          # it really was the un-instrumented kernel
    for e in range(n_envs):
        assert np.array_equal(sim.holdings(e), hold_i[e]), e
    sim.close()
    return st_i


def test_sparse_zi_1000_philox_production_kernel_equals_oracle():
    st = _philox_vs_oracle(lambda **kw: sparse_zi_config(1000, **{"draw_log_cap": 420000, **kw}), 64, 777000, 5)
    assert st["messages"].min() > 170000 and len(set(st["pop_hash"].tolist())) == 64


def test_sparse_zi_100_philox_production_kernel_equals_oracle():
    _philox_vs_oracle(lambda **kw: sparse_zi_config(100, **{"draw_log_cap": 60000, **kw}), 96, 31000, 5)


@pytest.mark.parametrize("pov", [False, True])
def test_rmsc03_philox_production_kernel_equals_oracle(pov):
    _philox_vs_oracle(lambda **kw: rmsc03_config(pov_exec=pov, **{"draw_log_cap": 60000, **kw}), 64, 5150, 4, skip_flag=_lib.F_OBS_INVALID)


def test_rmsc01_philox_production_kernel_equals_oracle():
    """config/rmsc01.py population, 09:30 - 09:36, 48 Philox-seeded environments: HBL belief argmax, order-history log and market-maker ladder of the
    production kernel == the instrumented kernel == the oracle on the logged draws."""
    stop = (9 * 3600 + 36 * 60) * 10 ** 9
    _philox_vs_oracle(lambda **kw: rmsc01_config(stop_ns=stop, **{"draw_log_cap": 60000, **kw}), 48, 8800, 5)


def test_rmsc02_philox_production_kernel_equals_oracle():
    """config/rmsc02.py (subscriptions + latency matrix), midnight - 11:00, 48 Philox-seeded environments: production kernel == instrumented kernel == oracle on the logged draws."""
    stop = 11 * 3600 * 10 ** 9
    _philox_vs_oracle(lambda **kw: rmsc02_config(stop_ns=stop, **{"draw_log_cap": 100000, **kw}), 48, 9900, 5)


def test_tape_mode_uninstrumented_kernel_equals_oracle():
    seeds = [123456789, 1001, 7, 424242]
    oracles = [OracleSim(100, s, TRACE_ALL) for s in seeds]
    counts = [o.run() for o in oracles]
    sim = BatchedSim(sparse_zi_config(100, rng_mode=_lib.RNG_TAPE), len(seeds))          # trace_cap 0, hash_pops 0: abx_run_kernel<TAPE, CUBIC, false>
    sim.reset_tape(*oracle_tapes(oracles))
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e, o in enumerate(oracles):
        assert_env_equals_oracle(sim, e, o, counts[e], st, traces=False, hashed=False)
    o3 = OracleSim(3, 1001, TRACE_ALL)
    n3 = o3.run()
    sim = BatchedSim(rmsc03_config(rng_mode=_lib.RNG_TAPE), 2)
    sim.reset_tape(*oracle_tapes([o3, o3]))
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(2):
        assert_env_equals_oracle(sim, e, o3, n3, st, traces=False, hashed=False, holdings_cols=4)


def _wave_tape_parity(cfg, oracles, counts, lib_path, holdings_cols):
    n_t = len(oracles)
    sim = BatchedSim(cfg, WAVE, lib_path=lib_path)
    sim.reset_tape_shared(n_t, *oracle_tapes(oracles))
    sim.run()
    sim.finalize()
    st = sim.stats()
    exp_hash = np.array([o.pop_hash() for o in oracles], dtype=np.uint64)[np.arange(WAVE) % n_t]
    assert np.array_equal(st["messages"], np.array(counts, dtype=np.int64)[np.arange(WAVE) % n_t])
    assert (st["flags"] == _lib.F_DONE).all(), np.unique(st["flags"])
    assert np.array_equal(st["pop_hash"], exp_hash), int((st["pop_hash"] != exp_hash).sum())
    for e in list(range(0, WAVE, 61)) + [WAVE - 1]:
        assert_env_equals_oracle(sim, e, oracles[e % n_t], counts[e % n_t], st, traces=False, holdings_cols=holdings_cols)
    sim.close()


@pytest.mark.parametrize("lib_path", LIBS)
def test_full_wave_tape_parity_sparse_zi_1000(lib_path):
    """The benchmarked shape at the benchmark's occupancy (16 co-resident one-warp CTAs per SM on the shared-memory carve-up)."""
    oracles = [OracleSim(1000, 123456789 + 1000 * k, TRACE_ALL) for k in range(8)]
    counts = [o.run() for o in oracles]
    assert counts[0] == 185200
    _wave_tape_parity(sparse_zi_config(1000, rng_mode=_lib.RNG_TAPE, hash_pops=1), oracles, counts, lib_path, 5)


@pytest.mark.parametrize("lib_path", LIBS)
def test_full_wave_tape_parity_rmsc03(lib_path):
    oracles = [OracleSim(3, s, TRACE_ALL) for s in (123456789, 1001, 5, 6, 7, 8, 9, 10)]
    counts = [o.run() for o in oracles]
    _wave_tape_parity(rmsc03_config(rng_mode=_lib.RNG_TAPE, hash_pops=1), oracles, counts, lib_path, 4)


@pytest.mark.parametrize("lib_path", LIBS)
def test_full_wave_tape_parity_rmsc01(lib_path):
    from helpers import oracle_rmsc01
    stop = (9 * 3600 + 34 * 60) * 10 ** 9
    runs = [oracle_rmsc01(s, stop, TRACE_ALL) for s in (123456789, 1001, 5, 6, 7, 8, 9, 10)]
    _wave_tape_parity(rmsc01_config(rng_mode=_lib.RNG_TAPE, hash_pops=1, stop_ns=stop), [r[0] for r in runs], [r[1] for r in runs], lib_path, 5)


@pytest.mark.parametrize("lib_path", LIBS)
def test_full_wave_tape_parity_rmsc02(lib_path):
    from helpers import oracle_rmsc02
    stop = int(10.5 * 3600) * 10 ** 9
    runs = [oracle_rmsc02(s, stop, TRACE_ALL) for s in (123456789, 1001, 5, 6, 7, 8, 9, 10)]
    _wave_tape_parity(rmsc02_config(rng_mode=_lib.RNG_TAPE, hash_pops=1, stop_ns=stop), [r[0] for r in runs], [r[1] for r in runs], lib_path, 5)


DAYS = ("env_IBM_2003-01-14_s789.npz", "env_IBM_2003-01-15_s4242.npz", "ddqn_IBM_2003-01-16_s99_sell.npz")


@pytest.mark.parametrize("lib_path", LIBS)
def test_full_wave_abidesenv_parity(golden_dir, lib_path):
    """2 368 ABIDESEnv environments (three replayed days round robin, a different action sequence per day): every observation of every
    environment and every pop hash equal to the oracle of its day."""
    from marl_optimal_execution_b200.env import ABIDESEnv, env_config
    from oracle.oracle import OracleEnv
    days = [np.load(os.path.join(golden_dir, f))["stream"] for f in DAYS]
    env = ABIDESEnv(days, n_envs=WAVE, cfg=env_config(hash_pops=1), lib_path=lib_path)
    env.reset()
    orc = [OracleEnv(d) for d in days]
    rs = np.random.RandomState(11)
    day_of = np.arange(WAVE) % 3
    steps = 0
    while True:
        a = np.stack([np.array([rs.uniform(0, 0.04), rs.uniform(), rs.uniform()]) for _ in range(3)])
        obs, rew, done, _ = env.step(torch.from_numpy(a[day_of]).cuda())
        obs, done = obs.cpu().numpy(), done.cpu().numpy()
        ods = []
        for d, o in enumerate(orc):
            x, _, od, _ = o.step(a[d])
            ods.append(od)
            ref = np.zeros(9); ref[: len(x)] = x
            assert np.allclose(obs[d], ref, rtol=1e-9, atol=1e-12), (steps, d)
            assert (obs[day_of == d] == obs[d]).all() and (done[day_of == d] == od).all(), (steps, d)
        steps += 1
        if all(ods):
            break
    st = env.stats()
    assert steps == 761
    assert np.array_equal(st["pop_hash"], np.array([o.pop_hash() for o in orc], dtype=np.uint64)[day_of])
    assert np.array_equal(st["messages"], np.array([o.n_pops for o in orc], dtype=np.int64)[day_of]) and (st["flags"] == _lib.F_DONE).all()
    env.close()


@pytest.mark.parametrize("lib_path", LIBS)
def test_full_wave_ddqn_shape_parity(golden_dir, lib_path):
    from marl_optimal_execution_b200.env import DDQNExecutionEnv, dq_config
    from oracle.oracle import OracleDDQNEnv
    gs = [np.load(os.path.join(golden_dir, f)) for f in DAYS]
    days = [g["stream"] for g in gs]
    ms = gs[2]["mom_sizes"].astype(np.int32)
    env = DDQNExecutionEnv(days, n_envs=WAVE, cfg=dq_config(hash_pops=1), lib_path=lib_path)
    env.reset(mom_sizes=np.tile(ms, (WAVE, 1)))
    orc = [OracleDDQNEnv(d, ms) for d in days]
    day_of = np.arange(WAVE) % 3
    rs = np.random.RandomState(5)
    obs, trans, rew, done = env.step(None)
    outs = [o.step(0) for o in orc]
    ticks = 0
    while not all(o[3] for o in outs):
        for d in range(3):
            if not outs[d][3]:
                assert np.allclose(obs[d], outs[d][0], rtol=1e-9, atol=1e-12), (ticks, d)
            assert (obs[day_of == d] == obs[d]).all(), (ticks, d)
        a = rs.randint(0, 24, 3).astype(np.int32)
        ot, tt, rt, dt = env.step(torch.from_numpy(a[day_of]).cuda())
        obs, trans, rew, done = ot.cpu().numpy(), tt.cpu().numpy(), rt.cpu().numpy(), dt.cpu().numpy()
        outs = [o.step(int(a[d])) if not outs[d][3] else outs[d] for d, o in enumerate(orc)]
        for d in range(3):
            assert np.isclose(rew[d], outs[d][2], rtol=1e-9, atol=1e-12) or outs[d][3], (ticks, d)
        ticks += 1
    st = env.stats()
    assert ticks >= 660
    assert np.array_equal(st["pop_hash"], np.array([o.pop_hash() for o in orc], dtype=np.uint64)[day_of])
    assert np.array_equal(st["messages"], np.array([o.n_pops for o in orc], dtype=np.int64)[day_of]) and (st["flags"] & _lib.F_ERROR_MASK == 0).all()
    env.close()
