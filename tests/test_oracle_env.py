"""Pin the oracle's ABIDESEnv restatement (GymKernel + MarketReplayAgent + DummyRLExecutionAgent + ABIDESEnvMetrics +
modifyOrder/history) against an episode recorded from the live reference (tools/record_reference_env.py)."""
import os

import numpy as np

from oracle.oracle import OracleEnv, TRACE_ALL


import pytest

EPISODES = [("env_IBM_2003-01-14_s789.npz", 144099), ("env_IBM_2003-01-15_s4242.npz", 165100)]


@pytest.mark.parametrize("fixture,n_pops", EPISODES)
def test_env_episode_matches_reference(golden_dir, fixture, n_pops):
    g = np.load(os.path.join(golden_dir, fixture))
    env = OracleEnv(g["stream"], quantity=1e5, order_level=2, trace=TRACE_ALL)
    obs, dones = [], []
    for a in g["actions"]:
        o, reward, done, info = env.step(a)
        assert reward is None and info is None                       # dummy_rl_execution_agent.py:325-352 returns None
        oo = np.full(9, np.nan)
        oo[: len(o)] = o
        obs.append(oo)
        dones.append(done)
        if done:
            break
    obs = np.array(obs)
    assert len(obs) == len(g["obs"]) == 761 and dones == list(g["done"])
    assert env.n_pops == int(g["n_pops"]) == n_pops
    assert np.array_equal(env.hash_ckpt(), g["pop_hash_ckpt"][:-1]) and env.pop_hash() == int(g["pop_hash_ckpt"][-1])
    assert env.note_hash() == int(g["note_hash"]) and env.snap_hash() == int(g["snap_hash"])
    for name in ("pops", "ops", "notes", "snaps"):
        a, b = env.trace(name), g[name + "_head"]
        assert np.array_equal(a[: len(b)], b), name
    # observations: fp64, 1e-6 relative (north star); measured 2e-15
    assert np.array_equal(np.isnan(obs), np.isnan(g["obs"]))
    err = np.nanmax(np.abs(obs - g["obs"]) / np.maximum(np.abs(g["obs"]), 1e-12))
    assert err < 1e-12, err                                           # north star: 1e-6 relative; measured < 1e-15
    f = env.final()
    assert np.array_equal(f[:4], g["rl_final"]) and np.array_equal(f[4:7], g["replay_final"])


def test_marketreplay_config_matches_reference(golden_dir):
    """config/marketreplay.py on GOOG 2012-06-21 (Exchange + MarketReplayAgent under Kernel.runner, stop 16:01): 193 264
    kernel messages, 49 482 book ops incl. 3 913 rows with ORDER_ID 0 (generated ids, util/order/Order.py:27)."""
    g = np.load(os.path.join(golden_dir, "mr_GOOG_2012-06-21.npz"))
    env = OracleEnv(g["stream"], order_level=0, trace=TRACE_ALL, stop_ns=(16 * 3600 + 60) * 10 ** 9)
    _, _, done, _ = env.step([0, 0, 0])
    assert done == 1 and env.n_pops == int(g["n_pops"]) == 193264
    assert np.array_equal(env.hash_ckpt(), g["pop_hash_ckpt"][:-1]) and env.pop_hash() == int(g["pop_hash_ckpt"][-1])
    assert env.note_hash() == int(g["note_hash"]) and env.snap_hash() == int(g["snap_hash"])
    for name in ("pops", "notes", "snaps"):
        b = g[name + "_head"]
        assert np.array_equal(env.trace(name)[: len(b)], b), name
    assert env.counter("max_bid_levels") == g["max_levels"][0] and env.counter("max_ask_levels") == g["max_levels"][1]
    assert env.counter("max_resting") == int(g["max_resting"])


def test_sample_orders_file_matches_reference(golden_dir):
    """data/sample_orders_file.csv (the 10-row L3 order stream) replayed by config/marketreplay.py in the live reference
    (tools/record_reference.py --orders-csv): 41 kernel messages; the first timestamp's two orders are replayed twice (second pass =
    MODIFY_ORDER), fills remove orders before their later CANCEL / MODIFY rows arrive, the last timestamp is never replayed.  Every pop,
    exchange message and book snapshot of the recording."""
    g = np.load(os.path.join(golden_dir, "mr_sample_orders_file.npz"))
    env = OracleEnv(g["stream"], order_level=0, trace=TRACE_ALL, stop_ns=(16 * 3600 + 60) * 10 ** 9)
    _, _, done, _ = env.step([0, 0, 0])
    assert done == 1 and env.n_pops == int(g["n_pops"]) == 41
    assert env.pop_hash() == int(g["pop_hash_ckpt"][-1]) and env.note_hash() == int(g["note_hash"]) and env.snap_hash() == int(g["snap_hash"])
    for name in ("pops", "notes", "snaps"):
        assert np.array_equal(env.trace(name), g[name]), name
    f = env.final()
    assert f[4] == g["holdings"][0, 1] == 0 and f[5] == g["holdings"][0, 2] == 0      # the replay agent trades with itself: flat, cash unchanged
