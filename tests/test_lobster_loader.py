"""Data format on the input side of the replay path: LOBSTER message files parsed like the reference's LOBSTEROrdersProcessor
(agent/examples/MarketReplayAgent.py:162-220), and the (ticker, date) constructor + gym spaces of the ABIDESEnv surface (ABIDESEnv.py:8-25)."""
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest

from helpers import build_emu
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.env import ABIDESEnv, Box, env_config, load_lobster_csv, lobster_message_path

REF_LOBSTER = "/root/reference/data/lobster"


@pytest.mark.skipif(not os.path.isdir(REF_LOBSTER), reason="reference tree not present")
@pytest.mark.parametrize("fixture,ticker,date,dated", [
    ("env_IBM_2003-01-14_s789.npz", "IBM", "2003-01-14", False), ("env_IBM_2003-01-15_s4242.npz", "IBM", "2003-01-15", False),
    ("ddqn_IBM_2003-01-16_s99_sell.npz", "IBM", "2003-01-16", False), ("mr_GOOG_2012-06-21.npz", "GOOG", "2012-06-21", True)])
def test_loader_equals_the_streams_the_live_reference_parsed(golden_dir, fixture, ticker, date, dated):
    """The fixtures hold orders_dict exactly as the reference's processor built it from these files (tools/record_reference*.py)."""
    want = np.load(os.path.join(golden_dir, fixture))["stream"]
    got = load_lobster_csv(lobster_message_path(ticker, date, REF_LOBSTER, 1, dated_folder=dated))
    assert got.dtype == np.int64 and np.array_equal(got, want)


@pytest.mark.skipif(not os.path.isdir(REF_LOBSTER), reason="reference tree not present")
@pytest.mark.parametrize("rel,date", [("LOBSTER_SampleFile_AMZN_2012-06-21_1/AMZN_2012-06-21", "2012-06-21")])
def test_loader_equals_the_live_reference_processor(tmp_path, rel, date):
    """A day no fixture holds: the unmodified LOBSTEROrdersProcessor run here (tools/reference_lobster_parse.py) vs load_lobster_csv
    (CSCO 2003-01-13, 88 315 rows, was checked the same way by hand; one day keeps the CPU suite short)."""
    csv = os.path.join(REF_LOBSTER, rel + "_34200000_57600000_message_1.csv")
    out = str(tmp_path / "ref.npy")
    subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "reference_lobster_parse.py"), csv, date, out],
                   check=True, capture_output=True, timeout=300)
    want = np.load(out)
    assert len(want) > 50000 and np.array_equal(load_lobster_csv(csv), want)


def test_empty_day_fails_like_the_reference(tmp_path):
    """An empty message file (the reference ships MSFT 2003-01-20, a market holiday): LOBSTEROrdersProcessor raises IndexError at
    wakeup_times[0] (MarketReplayAgent.py:177); so does the (ticker, date) constructor."""
    _write_day(tmp_path, "XYZ", "2020-01-01", [])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert load_lobster_csv(lobster_message_path("XYZ", "2020-01-01", str(tmp_path))).shape == (0, 5)
        with pytest.raises(IndexError):
            ABIDESEnv.from_lobster("XYZ", "2020-01-01", data_root=str(tmp_path), lib_path=build_emu())


def _write_day(root, ticker, date, lines):
    path = lobster_message_path(ticker, date, str(root))
    os.makedirs(os.path.dirname(path))
    with open(path, "w") as f:
        f.write("\n".join(lines) + ("\n" if lines else ""))
    return path


DAY = ["34199.999999999,1,5000011,100,1000000,1",          # before 09:30: dropped (MarketReplayAgent.py:212)
       "34200.017459617,1,5000012,100,1181599,1",          # PRICE 118.1599 $ -> int(1181599 / 100) = 11815 cents (truncation, :210-211)
       "34200.5,1,5000013,200,1190000,-1",
       "34200.5,3,5000012,100,1181599,1",                  # same timestamp: file order inside the group; event type is not looked at (:198-202)
       "34201.000000001,1,0,50,1185000,-1",           # ORDER_ID 0 (the exchange generates an id, util/order/Order.py:27)
       "57599.999999999,1,5000014,10,1180000,1",
       "57600.0,1,5000015,10,1180000,1"]                   # 16:00:00 itself: dropped (t < mkt_close)


def test_loader_rules_on_a_synthetic_day(tmp_path):
    path = _write_day(tmp_path, "XYZ", "2020-01-02", DAY)
    s = load_lobster_csv(path)
    assert s.tolist() == [[34200017459617, 5000012, 11815, 100, 1], [34200500000000, 5000013, 11900, 200, 0], [34200500000000, 5000012, 11815, 100, 1],
                          [34201000000001, 0, 11850, 50, 0], [57599999999999, 5000014, 11800, 10, 1]]


def test_ticker_date_constructor_and_spaces(tmp_path):
    """ABIDESEnv(ticker, date) of the reference -> ABIDESEnv.from_lobster(ticker, date, n_envs); action_space / observation_space as
    ABIDESEnv.py:18-25 builds them (actions in [0,1]^3; the observation Box has 10 zero bounds although 9 values come back)."""
    emu = build_emu()
    _write_day(tmp_path, "XYZ", "2020-01-02", DAY)
    L = _lib.load(emu)
    env = ABIDESEnv.from_lobster("XYZ", "2020-01-02", n_envs=2, data_root=str(tmp_path), cfg=env_config(L), lib_path=emu)
    assert (env.ticker, env.date, env.n_days) == ("XYZ", "2020-01-02", 1)
    assert isinstance(env.action_space, Box) and env.action_space.shape == (3,)
    assert env.action_space.low.tolist() == [0, 0, 0] and env.action_space.high.tolist() == [1, 1, 1]
    assert env.observation_space.shape == (10,) and not env.observation_space.high.any()
    a = env.action_space.sample(np.random.RandomState(1))
    assert env.action_space.contains(a) and not env.action_space.contains(a + 2)
    assert env.reset() is None
    obs, rew, done, info = env.step(np.stack([a, a]).astype(np.float64))
    assert obs.shape == (2, 9) and info is None and (env.stats()["flags"] & _lib.F_ERROR_MASK == 0).all()
    two = ABIDESEnv.from_lobster("XYZ", ["2020-01-02", "2020-01-02"], n_envs=2, data_root=str(tmp_path), cfg=env_config(L), lib_path=emu)
    assert two.n_days == 2
