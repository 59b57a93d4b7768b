"""CPU CI of the product's DDQN-execution-shape logic (abx_core.cuh compiled as plain C++ by tests/emu; a test tool, never a
fallback): the recorded reference runs of config/execution/marketreplay/execution_marketreplay_ddqn.py must be reproduced tick by
tick -- event order, exchange messages, book snapshots, observations, experience tuples, rewards, holdings."""
import os

import numpy as np
import pytest

from helpers import build_emu
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.env import DDQNExecutionEnv, dq_config
from oracle.oracle import OracleDDQNEnv, TRACE_ALL

FIXTURES = ["ddqn_IBM_2003-01-14_s4242.npz", "ddqn_IBM_2003-01-16_s99_sell.npz"]


@pytest.fixture(scope="module")
def emu():
    return build_emu()


def run_episode(env, o, g, n_envs):
    acts = g["actions"]
    obs, trans, rew, done = env.step(None)
    oo, otr, orw, od = o.step(0)
    k, total = 0, 0.0
    while not done[0]:
        assert not od and np.allclose(obs[0], oo, rtol=1e-12, atol=0), k
        assert np.allclose(obs[0, :6], g["observation"][k], rtol=1e-6, atol=1e-12), k        # vs the reference recording: 1e-6 relative, fp64
        a = np.full(n_envs, int(acts[k]), dtype=np.int32)
        if n_envs > 1:
            a[1] = (int(acts[k]) + 7) % 24                                                     # env 1 follows another policy: must diverge
        obs, trans, rew, done = env.step(a)
        oo, otr, orw, od = o.step(int(acts[k]))
        assert np.array_equal(np.nan_to_num(trans[0], nan=-7.0), np.nan_to_num(otr, nan=-7.0)), (k, trans[0], otr)
        assert rew[0] == orw, k
        total += rew[0]
        k += 1
    assert od and k == len(acts) == len(g["experience"])
    return k, total


@pytest.mark.parametrize("fixture", FIXTURES)
def test_ddqn_episode_matches_oracle_and_reference(emu, golden_dir, fixture):
    g = np.load(os.path.join(golden_dir, fixture))
    L = _lib.load(emu)
    is_buy = int(g["is_buy"]) if "is_buy" in g.files else 1
    env = DDQNExecutionEnv(g["stream"], n_envs=2, cfg=dq_config(L, is_buy=is_buy, trace_cap=420000, hash_pops=1), lib_path=emu)
    env.reset(mom_sizes=np.tile(g["mom_sizes"].astype(np.int32), (2, 1)))
    o = OracleDDQNEnv(g["stream"], g["mom_sizes"], is_buy=bool(is_buy), trace=TRACE_ALL)
    k, total = run_episode(env, o, g, 2)
    st = env.stats()
    assert int(st["messages"][0]) == o.n_pops == int(g["n_pops"]) and int(st["flags"][0]) == _lib.F_DONE and o.error() == 0
    assert int(st["pop_hash"][0]) == o.pop_hash() == int(g["pop_hash_ckpt"][-1])
    p, nt, sn = env.split_trace(0)
    assert np.array_equal(p, o.trace("pops")) and np.array_equal(nt, o.trace("notes")) and np.array_equal(sn, o.trace("snaps"))
    assert np.array_equal(p[: len(g["pops_head"])], g["pops_head"]) and np.array_equal(nt[: len(g["notes_head"])], g["notes_head"])
    assert abs(total - float(g["step_reward_hist"].sum())) < 1e-6 * abs(total)
    hold, ex = env.holdings(0)
    assert np.array_equal(hold[:, :4], g["holdings"][:, :4]) and np.array_equal(hold[:, :4], o.holdings()[:, :4])
    assert np.array_equal(hold[-2:, 4], g["holdings"][-2:, 4])                                # open orders of the two execution agents
    assert ex[0, 0] == g["twap_final"][0] and ex[0, 1] == g["twap_final"][1] and ex[0, 2] == g["twap_final"][2]
    assert ex[1, 0] == g["ddqn_final"][0] and ex[1, 1] == g["ddqn_final"][3] and ex[1, 2] == g["ddqn_final"][4] and ex[1, 4] == g["ddqn_final"][2]
    assert int(st["flags"][1]) == _lib.F_DONE and int(st["pop_hash"][1]) != int(st["pop_hash"][0])


def test_oracle_reproduces_reference_recording(golden_dir):
    """The oracle alone against everything the recorder kept of the live reference run."""
    for fixture in FIXTURES:
        g = np.load(os.path.join(golden_dir, fixture))
        is_buy = int(g["is_buy"]) if "is_buy" in g.files else 1
        o = OracleDDQNEnv(g["stream"], g["mom_sizes"], is_buy=bool(is_buy), trace=TRACE_ALL)
        out, tr, r, done = o.step(0)
        k = 0
        while not done:
            assert np.allclose(out[:6], g["observation"][k], rtol=1e-12, atol=1e-15)
            out, tr, r, done = o.step(int(g["actions"][k]))
            k += 1
        assert o.n_pops == int(g["n_pops"]) and o.pop_hash() == int(g["pop_hash_ckpt"][-1]) and o.error() == 0
        assert o.note_hash() == int(g["note_hash"]) and o.snap_hash() == int(g["snap_hash"])
        ck = o.hash_ckpt()
        assert np.array_equal(ck[: len(g["pop_hash_ckpt"]) - 1], g["pop_hash_ckpt"][:-1])
        ex, ge = o.series("experience"), g["experience"]
        assert np.array_equal(np.nan_to_num(ex, nan=-7.0), np.nan_to_num(ge, nan=-7.0))
        assert np.array_equal(o.series("price_path"), g["price_path"]) and np.array_equal(o.series("action_hist"), g["action_hist"])
        assert np.array_equal(o.series("step_reward_hist"), g["step_reward_hist"])
        assert np.array_equal(o.holdings()[:, :4], g["holdings"][:, :4])
        ops = o.trace("ops")
        assert np.array_equal(ops[ops[:, 2] >= 9], g["rl_ops"])


def test_closed_loop_with_a_real_network_oracle_side(golden_dir):
    """A third recorded reference run (IBM 2003-01-15) in which the agent's Keras model was replaced by a real fp32 network (numpy stand-in
    for util/model/QNets.py, weights in the fixture): the oracle environment driven by the same network arithmetic reproduces the reference's
    actions, event order and rewards -- choose_action (np.argmax of predict, :362-364) closed over the environment."""
    from marl_optimal_execution_b200.qnet import DEFAULT_DIMS, unpack_params
    g = np.load(os.path.join(golden_dir, "ddqn_mlp_IBM_2003-01-15_s31.npz"))
    stream = np.load(os.path.join(golden_dir, str(g["stream_fixture"])))["stream"]
    layers = unpack_params(g["mlp_params"], DEFAULT_DIMS)

    def policy(s):
        h = np.asarray(s, dtype=np.float32)[None, :]
        for i, (w, b) in enumerate(layers):
            h = h @ w.T + b
            if i + 1 < len(layers):
                h = np.maximum(h, np.float32(0))
        return h[0]

    o = OracleDDQNEnv(stream, g["mom_sizes"])
    out, tr, r, done = o.step(0)
    k, total = 0, 0.0
    while not done:
        q = policy(out[6:8])
        assert np.allclose(q, g["mlp_q"][k], rtol=1e-5, atol=1e-5) and int(np.argmax(q)) == int(g["actions"][k]), k
        out, tr, r, done = o.step(int(np.argmax(q)))
        total += r
        k += 1
    assert k == 660 and o.n_pops == int(g["n_pops"]) and o.pop_hash() == int(g["pop_hash_ckpt"][-1]) and o.error() == 0
    assert o.note_hash() == int(g["note_hash"]) and o.snap_hash() == int(g["snap_hash"])
    assert abs(total - float(g["step_reward_hist"].sum())) < 1e-9 * abs(total)


VWAP_FIXTURE = "ddqn_vwap_IBM_2003-01-15_s77.npz"


def load_vwap(golden_dir):
    g = np.load(os.path.join(golden_dir, VWAP_FIXTURE))
    stream = np.load(os.path.join(golden_dir, str(g["stream_fixture"])))["stream"]
    return g, stream


def test_vwap_schedule_helpers_equal_the_reference_agents(golden_dir):
    """`vwap_schedule(synthetic_volume_profile(n), quantity)` == the schedule the reference's VWAPExecutionAgent built from the same profile
    (tools/record_reference_ddqn.py --vwap stores both): 660 bins, 9 of them zero shares, total 500 003."""
    from marl_optimal_execution_b200.env import synthetic_volume_profile, vwap_schedule
    g, _ = load_vwap(golden_dir)
    w = synthetic_volume_profile(660)
    assert np.array_equal(w, g["vwap_profile"])
    sch = vwap_schedule(w, 500000)
    assert np.array_equal(sch, g["vwap_schedule"]) and int(sch.sum()) == 500003 and int((sch == 0).sum()) == 9


def test_oracle_reproduces_the_vwap_recording(golden_dir):
    """The config's baseline execution agent replaced by the reference's VWAPExecutionAgent (recorder --vwap): the oracle with the same per-bin schedule
    reproduces 212 827 pops, every exchange message and book snapshot, every book operation of the two execution agents (zero-quantity bins consume an
    order id and send nothing), the DDQN agent's observations, experience and rewards, and the holdings."""
    g, stream = load_vwap(golden_dir)
    o = OracleDDQNEnv(stream, g["mom_sizes"], is_buy=True, trace=TRACE_ALL)
    o.set_schedule(0, g["vwap_schedule"])
    out, tr, r, done = o.step(0)
    k = 0
    while not done:
        assert np.allclose(out[:6], g["observation"][k], rtol=1e-12, atol=1e-15)
        out, tr, r, done = o.step(int(g["actions"][k]))
        k += 1
    assert o.n_pops == int(g["n_pops"]) == 212827 and o.pop_hash() == int(g["pop_hash_ckpt"][-1]) and o.error() == 0
    assert o.note_hash() == int(g["note_hash"]) and o.snap_hash() == int(g["snap_hash"])
    ops = o.trace("ops")
    assert np.array_equal(ops[ops[:, 2] >= 9], g["rl_ops"])
    assert np.array_equal(o.series("step_reward_hist"), g["step_reward_hist"]) and np.array_equal(o.holdings()[:, :4], g["holdings"][:, :4])
    fin = o.exec_final(0) if hasattr(o, "exec_final") else None
    assert fin is None or (fin[0] == g["twap_final"][0] and fin[2] == g["twap_final"][2])


def test_vwap_episode_matches_oracle_and_reference(emu, golden_dir):
    """The product logic with `set_schedule` (abx_dq_set_schedule): the VWAP agent's episode tick by tick against the oracle and the recording."""
    g, stream = load_vwap(golden_dir)
    L = _lib.load(emu)
    env = DDQNExecutionEnv(stream, n_envs=2, cfg=dq_config(L, is_buy=1, trace_cap=460000, hash_pops=1), lib_path=emu)
    env.set_schedule(0, g["vwap_schedule"])
    env.reset(mom_sizes=np.tile(g["mom_sizes"].astype(np.int32), (2, 1)))
    o = OracleDDQNEnv(stream, g["mom_sizes"], is_buy=True, trace=TRACE_ALL)
    o.set_schedule(0, g["vwap_schedule"])
    k, total = run_episode(env, o, g, 2)
    st = env.stats()
    assert int(st["messages"][0]) == o.n_pops == int(g["n_pops"]) and int(st["flags"][0]) == _lib.F_DONE and o.error() == 0
    assert int(st["pop_hash"][0]) == o.pop_hash() == int(g["pop_hash_ckpt"][-1])
    p, nt, sn = env.split_trace(0)
    assert np.array_equal(p, o.trace("pops")) and np.array_equal(nt, o.trace("notes")) and np.array_equal(sn, o.trace("snaps"))
    hold, ex = env.holdings(0)
    assert np.array_equal(hold[:, :4], g["holdings"][:, :4])
    assert ex[0, 0] == g["twap_final"][0] and ex[0, 1] == g["twap_final"][1] and ex[0, 2] == g["twap_final"][2]
