"""GPU parity of the ABIDESEnv path through the C ABI: a recorded reference episode (761 steps, 144 099 kernel
messages) must reproduce bit-exactly in event order, exchange messages and book snapshots, and within 1e-6 relative
(fp64) in the observations."""
import os

import numpy as np
import pytest
import torch

from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.env import ABIDESEnv, env_config
from oracle.oracle import OracleEnv, TRACE_ALL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fixture", ["env_IBM_2003-01-14_s789.npz", "env_IBM_2003-01-15_s4242.npz"])
def test_episode_matches_reference_recording_and_oracle(golden_dir, fixture):
    g = np.load(os.path.join(golden_dir, fixture))
    env = ABIDESEnv(g["stream"], n_envs=3, cfg=env_config(trace_cap=300000, hash_pops=1))
    env.reset()
    o = OracleEnv(g["stream"], trace=TRACE_ALL)
    for k, a in enumerate(g["actions"]):
        acts = np.stack([a, a, a * np.array([0.25, 1.0, 1.0])])
        if k % 2:                                                        # alternate host-buffer and device-tensor entry points
            obs, rew, done, _ = env.step(acts)
        else:
            ot, rt, dt, _ = env.step(torch.from_numpy(acts).cuda())
            obs, rew, done = ot.cpu().numpy(), rt.cpu().numpy(), dt.cpu().numpy()
        oo, _, od, _ = o.step(a)
        ref = np.nan_to_num(g["obs"][k], nan=0.0)
        assert int(done[0]) == od == int(g["done"][k]) and rew[0] == 0.0, k
        assert np.allclose(obs[0], ref, rtol=1e-6, atol=1e-12), (k, obs[0], ref)
        assert np.array_equal(obs[0], obs[1]) and done[0] == done[1]
    st = env.stats()
    assert int(st["messages"][0]) == o.n_pops == int(g["n_pops"])
    assert (st["flags"] == _lib.F_DONE).all(), st["flags"]
    assert int(st["pop_hash"][0]) == int(st["pop_hash"][1]) == o.pop_hash() == int(g["pop_hash_ckpt"][-1])
    assert int(st["pop_hash"][2]) != int(st["pop_hash"][0])
    p, nt, sn = env.split_trace(0)
    for name, a_, b_ in (("pops", p, o.trace("pops")), ("notes", nt, o.trace("notes")), ("snaps", sn, o.trace("snaps"))):
        assert a_.shape == b_.shape, (name, a_.shape, b_.shape)
        d = np.nonzero((a_ != b_).any(axis=1))[0]
        assert len(d) == 0, (name, int(d[0]), a_[d[0]], b_[d[0]])
    assert np.array_equal(p[: len(g["pops_head"])], g["pops_head"]) and np.array_equal(nt[: len(g["notes_head"])], g["notes_head"])


def test_batch_of_identical_envs_is_deterministic(golden_dir):
    g = np.load(os.path.join(golden_dir, "env_IBM_2003-01-14_s789.npz"))
    n = 256
    env = ABIDESEnv(g["stream"], n_envs=n, cfg=env_config(hash_pops=1))
    env.reset()
    acts = torch.from_numpy(np.tile(g["actions"][:40, None, :], (1, n, 1))).cuda()
    for k in range(40):
        obs, _, done, _ = env.step(acts[k])
    st = env.stats()
    assert len(set(int(h) for h in st["pop_hash"])) == 1 and (st["flags"] == 0).all()
    assert torch.equal(obs[0], obs[-1]) and int(done.sum()) == 0


def test_marketreplay_config_on_gpu(golden_dir):
    """BASELINE.json configs[4]: config/marketreplay.py replaying a LOBSTER order stream through the GPU books --
    fills, snapshots and event order bit-exact against the reference recording (via the pinned oracle)."""
    g = np.load(os.path.join(golden_dir, "mr_GOOG_2012-06-21.npz"))
    stop = (16 * 3600 + 60) * 10 ** 9
    cfg = env_config(order_level=0, stop_ns=stop, queue_cap=256, level_cap=1024, trace_cap=400000, hash_pops=1)
    env = ABIDESEnv(g["stream"], n_envs=2, cfg=cfg)
    env.reset()
    o = OracleEnv(g["stream"], order_level=0, trace=TRACE_ALL, stop_ns=stop)
    _, _, done, _ = env.step(np.zeros((2, 3)))
    o.step([0, 0, 0])
    st = env.stats()
    assert done.tolist() == [1, 1] and (st["messages"] == 193264).all() and (st["flags"] == _lib.F_DONE).all()
    assert (st["pop_hash"] == np.uint64(int(g["pop_hash_ckpt"][-1]))).all()
    p, nt, sn = env.split_trace(1)
    assert np.array_equal(p, o.trace("pops")) and np.array_equal(nt, o.trace("notes")) and np.array_equal(sn, o.trace("snaps"))
    assert np.array_equal(nt[: len(g["notes_head"])], g["notes_head"]) and np.array_equal(sn[: len(g["snaps_head"])], g["snaps_head"])
    assert int(st["max_queue"][0]) == o.counter("max_queue")


def test_sample_orders_file_on_gpu(golden_dir):
    """data/sample_orders_file.csv replayed through the GPU books (config/marketreplay.py shape): every pop, exchange message and book
    snapshot of the live reference's own run (tools/record_reference.py --orders-csv), in three environments of one batch."""
    g = np.load(os.path.join(golden_dir, "mr_sample_orders_file.npz"))
    cfg = env_config(order_level=0, stop_ns=(16 * 3600 + 60) * 10 ** 9, queue_cap=64, level_cap=64, trace_cap=4096, hash_pops=1)
    env = ABIDESEnv(g["stream"], n_envs=3, cfg=cfg)
    env.reset()
    _, _, done, _ = env.step(np.zeros((3, 3)))
    st = env.stats()
    assert done.tolist() == [1, 1, 1] and (st["messages"] == 41).all() and (st["flags"] == _lib.F_DONE).all()
    assert (st["pop_hash"] == np.uint64(int(g["pop_hash_ckpt"][-1]))).all()
    for e in (0, 2):
        p, nt, sn = env.split_trace(e)
        assert np.array_equal(p, g["pops"]) and np.array_equal(nt, g["notes"]) and np.array_equal(sn, g["snaps"])


def test_several_days_in_one_batch(golden_dir):
    """Environment e replays day e % n_days (abx_env_create_days / abx_dq_create_days): three IBM days side by side, each environment
    bit-exact against the oracle of its own day; ABIDESEnv and the DDQN execution shape."""
    from marl_optimal_execution_b200.env import DDQNExecutionEnv, dq_config
    from oracle.oracle import OracleDDQNEnv
    gs = [np.load(os.path.join(golden_dir, f)) for f in ("env_IBM_2003-01-14_s789.npz", "env_IBM_2003-01-15_s4242.npz", "ddqn_IBM_2003-01-16_s99_sell.npz")]
    days = [g["stream"] for g in gs]
    env = ABIDESEnv(days, n_envs=6, cfg=env_config(hash_pops=1))
    env.reset()
    orc = [OracleEnv(d) for d in days]
    rs = np.random.RandomState(0)
    for k in range(150):
        a = np.array([rs.uniform(0, 0.04), rs.uniform(), rs.uniform()])
        obs, rew, done, _ = env.step(np.tile(a, (6, 1)))
        for d, o in enumerate(orc):
            x, _, _, _ = o.step(a)
            assert np.allclose(obs[d][: len(x)], x, rtol=1e-9, atol=1e-12) and np.array_equal(obs[d], obs[d + 3]), (k, d)
    st = env.stats()
    assert [int(h) for h in st["pop_hash"]] == [o.pop_hash() for o in orc] * 2 and (st["flags"] & _lib.F_ERROR_MASK == 0).all()
    ms = gs[2]["mom_sizes"].astype(np.int32)
    denv = DDQNExecutionEnv(days, n_envs=3, cfg=dq_config(hash_pops=1))
    denv.reset(mom_sizes=np.tile(ms, (3, 1)))
    dorc = [OracleDDQNEnv(d, ms) for d in days]
    obs, trans, rew, done = denv.step(None)
    for o in dorc:
        o.step(0)
    for k in range(60):
        a = int(rs.randint(0, 24))
        obs, trans, rew, done = denv.step(np.full(3, a, dtype=np.int32))
        for d, o in enumerate(dorc):
            oo, otr, orw, od = o.step(a)
            assert np.allclose(obs[d], oo, rtol=1e-9, atol=1e-12) and np.isclose(rew[d], orw, rtol=1e-9, atol=1e-12), (k, d)
    dst = denv.stats()
    assert [int(h) for h in dst["pop_hash"]] == [o.pop_hash() for o in dorc] and (dst["flags"] & _lib.F_ERROR_MASK == 0).all()
