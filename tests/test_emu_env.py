"""CPU CI of the product's ABIDESEnv logic (abx_core.cuh compiled as plain C++ by tests/emu; a test tool, never a
fallback): a full recorded episode must reproduce the oracle -- and therefore the reference -- step by step."""
import os

import numpy as np
import pytest

from helpers import build_emu
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.env import ABIDESEnv, env_config
from oracle.oracle import OracleEnv, TRACE_ALL


@pytest.fixture(scope="module")
def emu():
    return build_emu()


@pytest.mark.parametrize("fixture", ["env_IBM_2003-01-14_s789.npz", "env_IBM_2003-01-15_s4242.npz"])
def test_episode_matches_oracle_and_reference(emu, golden_dir, fixture):
    g = np.load(os.path.join(golden_dir, fixture))
    L = _lib.load(emu)
    env = ABIDESEnv(g["stream"], n_envs=2, cfg=env_config(L, trace_cap=300000, hash_pops=1), lib_path=emu)
    assert env.reset() is None and env.action_size == 3
    o = OracleEnv(g["stream"], trace=TRACE_ALL)
    for k, a in enumerate(g["actions"]):
        acts = np.stack([a, a * np.array([0.5, 1.0, 1.0])])            # env 1 takes half the size: diverges from env 0
        obs, rew, done, info = env.step(acts)
        oo, _, od, _ = o.step(a)
        ref = np.nan_to_num(g["obs"][k], nan=0.0)
        assert info is None and rew[0] == 0.0 and int(done[0]) == od == int(g["done"][k]), k
        assert np.allclose(obs[0], ref, rtol=1e-6, atol=1e-12), (k, obs[0], ref)   # vs the reference recording: 1e-6 relative, fp64
        oo9 = np.zeros(9)
        oo9[: len(oo)] = oo
        assert np.allclose(obs[0], oo9, rtol=1e-9, atol=1e-12), k
    st = env.stats()
    assert int(st["messages"][0]) == o.n_pops == int(g["n_pops"]) and int(st["flags"][0]) == _lib.F_DONE
    assert int(st["pop_hash"][0]) == o.pop_hash() == int(g["pop_hash_ckpt"][-1])
    p, nt, sn = env.split_trace(0)
    assert np.array_equal(p, o.trace("pops")) and np.array_equal(nt, o.trace("notes")) and np.array_equal(sn, o.trace("snaps"))
    assert int(st["flags"][1]) == _lib.F_DONE and int(st["pop_hash"][1]) != int(st["pop_hash"][0])


def test_generated_ids_skip_the_streams_explicit_ids(emu):
    from reset_cases import generated_ids_skip_explicit_ids
    generated_ids_skip_explicit_ids(emu)


def test_marketreplay_config_matches_oracle(emu, golden_dir):
    g = np.load(os.path.join(golden_dir, "mr_GOOG_2012-06-21.npz"))
    L = _lib.load(emu)
    stop = (16 * 3600 + 60) * 10 ** 9
    cfg = env_config(L, order_level=0, stop_ns=stop, queue_cap=256, level_cap=1024, trace_cap=400000, hash_pops=1)
    env = ABIDESEnv(g["stream"], n_envs=1, cfg=cfg, lib_path=emu)
    env.reset()
    o = OracleEnv(g["stream"], order_level=0, trace=TRACE_ALL, stop_ns=stop)
    _, _, done, _ = env.step(np.zeros((1, 3)))
    o.step([0, 0, 0])
    st = env.stats()[0]
    assert int(done[0]) == 1 and int(st["messages"]) == o.n_pops == 193264 and int(st["flags"]) == _lib.F_DONE
    assert int(st["pop_hash"]) == o.pop_hash() == int(g["pop_hash_ckpt"][-1])
    p, nt, sn = env.split_trace(0)
    assert np.array_equal(p, o.trace("pops")) and np.array_equal(nt, o.trace("notes")) and np.array_equal(sn, o.trace("snaps"))


def test_sample_orders_file_matches_reference_recording(emu, golden_dir):
    """The reference's 10-row L3 sample stream (data/sample_orders_file.csv) under config/marketreplay.py: all 41 messages of the live
    reference's run, its 17 exchange messages and 10 book snapshots."""
    g = np.load(os.path.join(golden_dir, "mr_sample_orders_file.npz"))
    L = _lib.load(emu)
    cfg = env_config(L, order_level=0, stop_ns=(16 * 3600 + 60) * 10 ** 9, queue_cap=64, level_cap=64, trace_cap=4096, hash_pops=1)
    env = ABIDESEnv(g["stream"], n_envs=1, cfg=cfg, lib_path=emu)
    env.reset()
    _, _, done, _ = env.step(np.zeros((1, 3)))
    st = env.stats()[0]
    assert int(done[0]) == 1 and int(st["messages"]) == int(g["n_pops"]) == 41 and int(st["flags"]) == _lib.F_DONE
    assert int(st["pop_hash"]) == int(g["pop_hash_ckpt"][-1])
    p, nt, sn = env.split_trace(0)
    assert np.array_equal(p, g["pops"]) and np.array_equal(nt, g["notes"]) and np.array_equal(sn, g["snaps"])


def test_several_days_in_one_batch(emu, golden_dir):
    """Environment e replays day e % n_days: two recorded reference episodes (IBM 2003-01-14 and 2003-01-15) side by side in one handle,
    each bit-exact against its own oracle; a third and fourth environment repeat the days."""
    ga = np.load(os.path.join(golden_dir, "env_IBM_2003-01-14_s789.npz"))
    gb = np.load(os.path.join(golden_dir, "env_IBM_2003-01-15_s4242.npz"))
    L = _lib.load(emu)
    env = ABIDESEnv([ga["stream"], gb["stream"]], n_envs=4, cfg=env_config(L, hash_pops=1), lib_path=emu)
    env.reset()
    oa, ob = OracleEnv(ga["stream"]), OracleEnv(gb["stream"])
    n = 120
    for k in range(n):
        acts = np.stack([ga["actions"][k], gb["actions"][k], ga["actions"][k], gb["actions"][k]])
        obs, rew, done, _ = env.step(acts)
        xa, _, _, _ = oa.step(ga["actions"][k]); xb, _, _, _ = ob.step(gb["actions"][k])
        assert np.allclose(obs[0][: len(xa)], xa, rtol=1e-9, atol=1e-12) and np.allclose(obs[1][: len(xb)], xb, rtol=1e-9, atol=1e-12), k
        assert np.allclose(obs[0, :9], np.nan_to_num(ga["obs"][k], nan=0.0), rtol=1e-6, atol=1e-12) and np.allclose(obs[1, :9], np.nan_to_num(gb["obs"][k], nan=0.0), rtol=1e-6, atol=1e-12)
    st = env.stats()
    assert int(st["pop_hash"][0]) == int(st["pop_hash"][2]) == oa.pop_hash() and int(st["pop_hash"][1]) == int(st["pop_hash"][3]) == ob.pop_hash()
    assert int(st["messages"][0]) == oa.n_pops and int(st["messages"][1]) == ob.n_pops and (st["flags"] & _lib.F_ERROR_MASK == 0).all()
