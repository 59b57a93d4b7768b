"""Per-environment reset, day rotation, auto-reset and order_level 1 on the CUDA kernels (the cases of tests/reset_cases.py), plus the masked reset
with a CUDA mask tensor at batch scale."""
import os

import numpy as np
import pytest
import torch

from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.env import ABIDESEnv, env_config
from reset_cases import DAYS, abidesenv_masked_reset_and_rotation, ddqn_auto_reset, generated_ids_skip_explicit_ids, order_level_one

pytestmark = pytest.mark.gpu


def test_abidesenv_masked_reset_and_day_rotation(golden_dir):
    abidesenv_masked_reset_and_rotation(golden_dir, n_steps=120)


def test_ddqn_shape_auto_reset_next_day(golden_dir):
    ddqn_auto_reset(golden_dir, ticks_after=60)


def test_order_level_one_actions(golden_dir):
    order_level_one(golden_dir, n_steps=200)


def test_generated_ids_skip_the_streams_explicit_ids():
    generated_ids_skip_explicit_ids()


def test_train_date_whose_ids_reach_into_the_generated_range(golden_dir):
    """IBM 2003-01-17 (one of the reference's nine train dates): its smallest LOBSTER ORDER_ID is 10 511, inside the range the day's generated ids run through
    (momentum + execution agents); the whole DDQN-shape day must match the oracle, which keeps the reference's used-id list."""
    from marl_optimal_execution_b200.env import DDQNExecutionEnv, dq_config
    from oracle.oracle import OracleDDQNEnv
    stream = np.load(os.path.join(golden_dir, "days", "IBM_2003-01-17.npz"))["stream"]
    ms = np.array([5, 3, 9, 2, 7, 4, 6], dtype=np.int32)
    env = DDQNExecutionEnv(stream, n_envs=2, cfg=dq_config(hash_pops=1))
    env.reset(mom_sizes=np.tile(ms, (2, 1)))
    o = OracleDDQNEnv(stream, ms)
    rs = np.random.RandomState(2)
    obs, trans, rew, done = env.step(None)
    out = o.step(0)
    while not done.all():
        assert np.allclose(obs[0], out[0], rtol=1e-9, atol=1e-12)
        a = int(rs.randint(0, 24))
        obs, trans, rew, done = env.step(np.full(2, a, dtype=np.int32))
        out = o.step(a)
    st = env.stats()
    assert (st["pop_hash"] == np.uint64(o.pop_hash())).all() and (st["messages"] == o.n_pops).all() and (st["flags"] & _lib.F_ERROR_MASK == 0).all()
    assert int(st["orders_allocated"][0]) > 5000                               # thousands of generated ids next to a stream whose explicit ids start at 10 511 (round 1 rejected this day at create time)
    env.close()


def test_auto_reset_keeps_a_batch_running_across_episodes(golden_dir):
    """512 ABIDESEnv environments, auto-reset on: every environment finishes its 761-step episode, restarts, and the second episode reproduces the
    first one's pop hash at the same step count (same day, same actions) -- nothing of the finished episode leaks into the next."""
    g = np.load(os.path.join(golden_dir, DAYS[0]))
    n = 512
    env = ABIDESEnv(g["stream"], n_envs=n, cfg=env_config(hash_pops=1))
    env.set_auto_reset(1)
    env.reset()
    acts = torch.from_numpy(np.tile(g["actions"][:, None, :], (1, n, 1))).cuda()
    hashes = []
    for ep in range(2):
        for k in range(len(acts)):
            obs, _, done, _ = env.step(acts[k])
            if k == 300:
                hashes.append(env.stats()["pop_hash"].copy())
        assert int(done.sum()) == n                                # the last step of the episode reports done; the reset follows it
    assert np.array_equal(hashes[0], hashes[1]) and len(set(hashes[0].tolist())) == 1
    assert (env.stats()["flags"] & _lib.F_ERROR_MASK == 0).all()
    env.close()
