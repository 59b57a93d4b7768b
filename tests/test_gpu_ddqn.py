"""GPU parity of the DDQN execution shape through the C ABI: two recorded reference runs of
config/execution/marketreplay/execution_marketreplay_ddqn.py (BUY and SELL, ~190 000 kernel messages, 660 decision ticks each)
must reproduce bit-exactly in event order, exchange messages, book snapshots, experience tuples and holdings, and within
1e-6 relative (fp64) in the observation features and rewards."""
import os

import numpy as np
import pytest
import torch

from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.env import DDQNExecutionEnv, dq_config
from oracle.oracle import OracleDDQNEnv, TRACE_ALL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fixture", ["ddqn_IBM_2003-01-14_s4242.npz", "ddqn_IBM_2003-01-16_s99_sell.npz"])
def test_ddqn_run_matches_reference_recording_and_oracle(golden_dir, fixture):
    g = np.load(os.path.join(golden_dir, fixture))
    is_buy = int(g["is_buy"])
    env = DDQNExecutionEnv(g["stream"], n_envs=3, cfg=dq_config(is_buy=is_buy, trace_cap=420000, hash_pops=1))
    env.reset(mom_sizes=np.tile(g["mom_sizes"].astype(np.int32), (3, 1)))
    o = OracleDDQNEnv(g["stream"], g["mom_sizes"], is_buy=bool(is_buy), trace=TRACE_ALL)
    acts = g["actions"]
    obs, trans, rew, done = env.step(None)
    oo, otr, orw, od = o.step(0)
    k, total = 0, 0.0
    while not done[0]:
        assert not od and np.allclose(obs[0], oo, rtol=1e-9, atol=1e-12), (k, obs[0], oo)
        assert np.allclose(obs[0, :6], g["observation"][k], rtol=1e-6, atol=1e-12), k         # vs the reference recording: 1e-6 relative, fp64
        assert np.array_equal(obs[0], obs[1])
        a = np.array([acts[k], acts[k], (acts[k] + 7) % 24], dtype=np.int32)                  # env 2 follows another policy
        if k % 2:                                                                             # alternate host-buffer and device-tensor entry points
            obs, trans, rew, done = env.step(a)
        else:
            ot, tt, rt, dt = env.step(torch.from_numpy(a).cuda())
            obs, trans, rew, done = ot.cpu().numpy(), tt.cpu().numpy(), rt.cpu().numpy(), dt.cpu().numpy()
        oo, otr, orw, od = o.step(int(acts[k]))
        assert np.array_equal(trans[0, :5], otr[:5]) and np.allclose(np.nan_to_num(trans[0, 5], nan=-7.0), np.nan_to_num(otr[5], nan=-7.0), rtol=1e-9), (k, trans[0], otr)
        assert np.isclose(rew[0], orw, rtol=1e-9, atol=1e-12), k
        total += rew[0]
        k += 1
    assert od and k == len(acts) == 660
    st = env.stats()
    assert int(st["messages"][0]) == o.n_pops == int(g["n_pops"]) and (st["flags"] == _lib.F_DONE).all(), st["flags"]
    assert int(st["pop_hash"][0]) == int(st["pop_hash"][1]) == o.pop_hash() == int(g["pop_hash_ckpt"][-1])
    assert int(st["pop_hash"][2]) != int(st["pop_hash"][0])
    p, nt, sn = env.split_trace(0)
    assert np.array_equal(p, o.trace("pops")) and np.array_equal(nt, o.trace("notes")) and np.array_equal(sn, o.trace("snaps"))
    assert abs(total - float(g["step_reward_hist"].sum())) < 1e-6 * abs(total)
    hold, ex = env.holdings(0)
    assert np.array_equal(hold[:, :4], g["holdings"][:, :4]) and np.array_equal(hold[-2:, 4], g["holdings"][-2:, 4])
    assert ex[0, 0] == g["twap_final"][0] and ex[0, 1] == g["twap_final"][1] and ex[0, 2] == g["twap_final"][2]
    assert ex[1, 0] == g["ddqn_final"][0] and ex[1, 1] == g["ddqn_final"][3] and ex[1, 2] == g["ddqn_final"][4] and ex[1, 4] == g["ddqn_final"][2]


def test_vwap_execution_agent_matches_reference_recording_and_oracle(golden_dir):
    """the config's baseline execution agent as the reference's VWAPExecutionAgent (tools/record_reference_ddqn.py --vwap): per-bin schedule
    through abx_dq_set_schedule; event order, exchange messages, snapshots, the agent's fills and the DDQN agent's rewards equal the oracle, which reproduces the
    live recording (tests/test_emu_ddqn.py::test_oracle_reproduces_the_vwap_recording)."""
    g = np.load(os.path.join(golden_dir, "ddqn_vwap_IBM_2003-01-15_s77.npz"))
    stream = np.load(os.path.join(golden_dir, str(g["stream_fixture"])))["stream"]
    env = DDQNExecutionEnv(stream, n_envs=2, cfg=dq_config(is_buy=1, trace_cap=460000, hash_pops=1))
    env.set_schedule(0, g["vwap_schedule"])
    env.reset(mom_sizes=np.tile(g["mom_sizes"].astype(np.int32), (2, 1)))
    o = OracleDDQNEnv(stream, g["mom_sizes"], is_buy=True, trace=TRACE_ALL)
    o.set_schedule(0, g["vwap_schedule"])
    acts = g["actions"]
    obs, trans, rew, done = env.step(None)
    oo, otr, orw, od = o.step(0)
    k = 0
    while not done[0]:
        assert not od and np.allclose(obs[0], oo, rtol=1e-9, atol=1e-12), k
        assert np.allclose(obs[0, :6], g["observation"][k], rtol=1e-6, atol=1e-12), k
        obs, trans, rew, done = env.step(np.full(2, int(acts[k]), dtype=np.int32))
        oo, otr, orw, od = o.step(int(acts[k]))
        assert np.isclose(rew[0], orw, rtol=1e-9, atol=1e-12), k
        k += 1
    assert od and k == 660
    st = env.stats()
    assert int(st["messages"][0]) == o.n_pops == int(g["n_pops"]) == 212827 and (st["flags"] == _lib.F_DONE).all(), st["flags"]
    assert int(st["pop_hash"][0]) == int(st["pop_hash"][1]) == o.pop_hash() == int(g["pop_hash_ckpt"][-1])
    p, nt, sn = env.split_trace(0)
    assert np.array_equal(p, o.trace("pops")) and np.array_equal(nt, o.trace("notes")) and np.array_equal(sn, o.trace("snaps"))
    hold, ex = env.holdings(0)
    assert np.array_equal(hold[:, :4], g["holdings"][:, :4])
    assert ex[0, 0] == g["twap_final"][0] and ex[0, 1] == g["twap_final"][1] and ex[0, 2] == g["twap_final"][2]


def test_ddqn_batch_is_deterministic_and_seed_dependent(golden_dir):
    """Philox-drawn MomentumAgent sizes: same seed -> identical environments, different seeds -> different runs; no error flags."""
    g = np.load(os.path.join(golden_dir, "ddqn_IBM_2003-01-14_s4242.npz"))
    n = 64
    env = DDQNExecutionEnv(g["stream"], n_envs=n, cfg=dq_config(hash_pops=1))
    seeds = np.arange(n, dtype=np.uint64) // 2 + 1000                      # pairs of equal seeds
    env.reset(seeds=seeds)
    rs = np.random.RandomState(3)
    obs, trans, rew, done = env.step(None)
    for k in range(40):
        a = np.repeat(rs.randint(0, 24, n // 2), 2).astype(np.int32)
        obs, trans, rew, done = env.step(torch.from_numpy(a).cuda())
    st = env.stats()
    assert (st["flags"] & _lib.F_ERROR_MASK == 0).all(), st["flags"]
    h = st["pop_hash"]
    assert (h[0::2] == h[1::2]).all() and len(set(h.tolist())) == n // 2
    assert not done.cpu().numpy().any() and (st["messages"] > 10000).all()
