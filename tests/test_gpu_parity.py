"""GPU parity tests (run on the B200 box, through the C ABI): RNG draws recorded by the oracle are replayed through
the CUDA simulator; event order, every outbound exchange message (fills, L1 replies), the book state after every
book op and the final holdings must equal the oracle's -- bit-exact (integer work).  The oracle itself is pinned to
the reference's recorded runs by the CPU suite (test_oracle_golden.py)."""
import os

import numpy as np
import pytest

from helpers import oracle_tapes
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.sim import BatchedSim, sparse_zi_config
from oracle.oracle import OracleSim, TRACE_ALL

pytestmark = pytest.mark.gpu


def _check_env(sim, e, o, n, st):
    assert int(st["messages"][e]) == n
    assert int(st["flags"][e]) == _lib.F_DONE, hex(int(st["flags"][e]))
    assert int(st["pop_hash"][e]) == o.pop_hash()
    p, nt, sn = sim.split_trace(e)
    for name, a, b in (("pops", p, o.trace("pops")), ("notes", nt, o.trace("notes")), ("snaps", sn, o.trace("snaps"))):
        assert a.shape == b.shape, (name, a.shape, b.shape)
        d = np.nonzero((a != b).any(axis=1))[0]
        assert len(d) == 0, (name, int(d[0]), a[d[0]], b[d[0]])
    assert np.array_equal(sim.holdings(e), o.holdings())
    assert int(st["limit_orders"][e]) == o.counter("limit") and int(st["cancels"][e]) == o.counter("cancel")
    assert int(st["fills"][e]) == o.counter("fills") and int(st["spread_queries"][e]) == o.counter("spread_queries")
    assert int(st["max_queue"][e]) == o.counter("max_queue") and int(st["uniq"][e]) == o.counter("uniq")
    l1 = o.book_l1()
    assert (int(st["best_bid"][e]), int(st["best_bid_qty"][e]), int(st["best_ask"][e]), int(st["best_ask_qty"][e]),
            int(st["last_trade"][e])) == tuple(int(x) for x in l1)
    assert int(st["fundamental"][e]) == o.fundamental()
    assert int(st["sum_shares"][e]) == 0 and int(st["sum_cash"][e]) == (o.n_agents - 1) * 10 ** 7


def test_z100_tape_replay_batch():
    seeds = [123456789, 1001, 7, 424242]
    oracles = [OracleSim(100, s, TRACE_ALL) for s in seeds]
    counts = [o.run() for o in oracles]
    cfg = sparse_zi_config(100, rng_mode=_lib.RNG_TAPE, trace_cap=60000, hash_pops=1)
    sim = BatchedSim(cfg, len(seeds))
    sim.reset_tape(*oracle_tapes(oracles))
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e, o in enumerate(oracles):
        _check_env(sim, e, o, counts[e], st)
    # depth-k snapshot accessor vs the trace's last snapshot row
    _, _, sn = sim.split_trace(0)
    bids = sim.book_snapshot(0, True, 3)
    assert [x for pq in bids for x in pq] == [int(v) for v in sn[-1][3:3 + 2 * len(bids)]]


def test_z1000_tape_replay_reproduces_reference_golden(golden_dir):
    """config/sparse_zi_1000.py seed 123456789 == the reference's tests/sparse_zi_1000.txt run (185 200 messages)."""
    o = OracleSim(1000, 123456789, TRACE_ALL)
    n = o.run()
    assert n == 185200
    cfg = sparse_zi_config(1000, rng_mode=_lib.RNG_TAPE, trace_cap=300000, hash_pops=1)
    sim = BatchedSim(cfg, 1)
    sim.reset_tape(*oracle_tapes([o]))
    sim.run()
    sim.finalize()
    st = sim.stats()
    _check_env(sim, 0, o, n, st)
    g = np.load(os.path.join(golden_dir, "z1000_s123456789.npz"))        # recorded from the live reference
    assert np.array_equal(sim.holdings(0), g["holdings"])
    assert int(st["pop_hash"][0]) == int(g["pop_hash_ckpt"][-1])
    # Kernel.runner's stdout contract (Kernel.py:321-343, TradingAgent.py:115-138): first line and the seven per-type means of the
    # reference's own capture tests/sparse_zi_1000.txt:22-32 (the CPU suite compares all 1 008 lines with that file)
    text = sim.kernel_summary(0, "JPM", elapsed_s=59.733625)
    assert text[0] == "Final holdings for ZI Agent 1 Type 1 [0 <= R <= 250, eta=1]: { JPM: 400, CASH: -30025600 }.  Marked to market: 9586000"
    assert text[1000] == "Event Queue elapsed: 0 days 00:00:59.733625, messages: 185200, messages per second: 3100.4"
    assert [int(ln.rsplit(": ", 1)[1]) for ln in text[1002:1009]] == [-19781, -19338, -18593, 30137, 641, 5634, 27887]


def test_sliced_runs_equal_single_run():
    o = OracleSim(100, 1001, TRACE_ALL)
    o.run()
    cfg = sparse_zi_config(100, rng_mode=_lib.RNG_TAPE, hash_pops=1)
    sim = BatchedSim(cfg, 1)
    sim.reset_tape(*oracle_tapes([o]))
    for q in range(0, 17 * 4):
        sim.run(q * 900 * 10 ** 9)
    sim.run()
    st = sim.stats()
    assert int(st["pop_hash"][0]) == o.pop_hash() and int(st["messages"][0]) == o.n_pops


def test_intermediate_state_matches_oracle():
    """Stop both at 11:00 and compare the live book + counters (not only end-of-day state)."""
    o = OracleSim(100, 123456789, TRACE_ALL)
    o.run()
    o2 = OracleSim(100, 123456789, 0)
    t = 11 * 3600 * 10 ** 9
    n2, done = o2.run_until(t)
    cfg = sparse_zi_config(100, rng_mode=_lib.RNG_TAPE, hash_pops=1)
    sim = BatchedSim(cfg, 1)
    sim.reset_tape(*oracle_tapes([o]))
    sim.run(t)
    st = sim.stats()
    assert int(st["messages"][0]) == n2 and not done and int(st["flags"][0]) == 0
    assert int(st["pop_hash"][0]) == o2.pop_hash()
    l1 = o2.book_l1()
    assert (int(st["best_bid"][0]), int(st["best_bid_qty"][0]), int(st["best_ask"][0]), int(st["best_ask_qty"][0]),
            int(st["last_trade"][0])) == tuple(int(x) for x in l1)


@pytest.mark.parametrize("seed", [123456789, 1001])
def test_rmsc03_tape_replay(seed):
    """config/rmsc03.py (BASELINE.json configs[2]): value + noise + momentum + POV market-maker population; replayed draws incl.
    the GLOBAL np.random stream; transacted-volume driven ladder sizes.  Bit-exact vs the oracle, which is pinned to
    recordings of the live reference (tests/test_oracle_golden.py::test_rmsc03_digest_bit_exact)."""
    from marl_optimal_execution_b200.sim import rmsc03_config
    o = OracleSim(3, seed, TRACE_ALL)
    n = o.run()
    cfg = rmsc03_config(rng_mode=_lib.RNG_TAPE, trace_cap=400000, hash_pops=1)
    sim = BatchedSim(cfg, 2)
    sim.reset_tape(*oracle_tapes([o, o]))
    sim.run()
    sim.finalize()
    st = sim.stats()
    assert (st["messages"] == n).all() and (st["flags"] == _lib.F_DONE).all(), (st["messages"], st["flags"])
    assert (st["pop_hash"] == np.uint64(o.pop_hash())).all()
    p, nt, sn = sim.split_trace(1)
    for name, a, b in (("pops", p, o.trace("pops")), ("notes", nt, o.trace("notes")), ("snaps", sn, o.trace("snaps"))):
        assert a.shape == b.shape, (name, a.shape, b.shape)
        d = np.nonzero((a != b).any(axis=1))[0]
        assert len(d) == 0, (name, int(d[0]), a[d[0]], b[d[0]])
    assert np.array_equal(sim.holdings(0)[:, :4], o.holdings()[:, :4])
    assert int(st["limit_orders"][0]) == o.counter("limit") and int(st["fills"][0]) == o.counter("fills")


@pytest.mark.parametrize("seed,stop_s,hist_cap", [(123456789, 15 * 60, 0), (20231, 4 * 60, 0), (123456789, 15 * 60, 256)])
def test_rmsc01_tape_replay(seed, stop_s, hist_cap):
    """config/rmsc01.py population: MarketMakerAgent, ZI, HeuristicBeliefLearningAgents served by QUERY_ORDER_STREAM, Momentum agents.
    Bit-exact pops, exchange messages, book snapshots and holdings vs the oracle, which is pinned to a live recording of the reference for seed
    123456789 up to 09:45:00 (tests/test_oracle_golden.py::test_rmsc01_full_trace_bit_exact)."""
    from helpers import assert_env_equals_oracle, oracle_rmsc01
    from marl_optimal_execution_b200.sim import rmsc01_config
    stop = (9 * 3600 + 30 * 60 + stop_s) * 10 ** 9
    o, n = oracle_rmsc01(seed, stop, TRACE_ALL)
    cfg = rmsc01_config(rng_mode=_lib.RNG_TAPE, trace_cap=400000, hash_pops=1, stop_ns=stop)
    if hist_cap:                                          # 256 scratch rows: wider price spans take the candidate form of the HBL argmax
        cfg.hbl_table_rows = hist_cap
    sim = BatchedSim(cfg, 2)
    sim.reset_tape(*oracle_tapes([o, o]))
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(2):
        assert_env_equals_oracle(sim, e, o, n, st)
    assert int(st["sum_shares"][0]) == 0 and int(st["sum_cash"][0]) == 100 * 10 ** 7


@pytest.mark.parametrize("fixture,seed", [("rmsc03_aggressive_s123456789.npz", 123456789), ("rmsc03_passive_s123456789.npz", 123456789), ("rmsc03_passive_limit_s1001.npz", 1001)])
def test_rmsc03_with_passive_or_aggressive_agent_tape_replay(golden_dir, fixture, seed):
    """PassiveAgent / AggressiveAgent (agent/execution/baselines/*.py) as exec_kind 1 / 2 of the rmsc03 population's execution-agent slot; full traces
    vs the oracle, which reproduces recordings of the reference with those agents appended (tests/test_oracle_golden.py::test_rmsc03_with_passive_or_aggressive_agent)."""
    import os
    from helpers import assert_env_equals_oracle
    from test_oracle_golden import exec_agent_config
    from marl_optimal_execution_b200.sim import rmsc03_config
    g = np.load(os.path.join(golden_dir, fixture))
    oc = exec_agent_config(g)
    o = OracleSim.from_config(oc, seed, TRACE_ALL)
    n = o.run()
    assert n == int(g["n_pops"])
    cfg = rmsc03_config(pov_exec=True, rng_mode=_lib.RNG_TAPE, trace_cap=400000, hash_pops=1, exec_kind=oc.exec_kind, pov_exec_start_ns=oc.pov_exec_start_ns,
                        pov_exec_quantity=oc.pov_exec_quantity, pov_exec_is_buy=oc.pov_exec_is_buy, exec_limit_price=oc.exec_limit_price)
    sim = BatchedSim(cfg, 2)
    sim.reset_tape(*oracle_tapes([o, o]))
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(2):
        assert_env_equals_oracle(sim, e, o, n, st, holdings_cols=4)


def test_rmsc02_whole_day_tape_replay():
    """config/rmsc02.py, the whole day (117 238 messages): MARKET_DATA subscriptions, subscription-mode market maker / momentum agents, latency matrix + noise.
    Full traces vs the oracle, which is pinned to a live recording of the reference for this seed (tests/test_oracle_golden.py::test_rmsc02_full_day_bit_exact)."""
    from helpers import assert_env_equals_oracle, oracle_rmsc02
    from marl_optimal_execution_b200.sim import rmsc02_config
    stop = 17 * 3600 * 10 ** 9
    o, n = oracle_rmsc02(123456789, stop, TRACE_ALL)
    assert n == 117238
    cfg = rmsc02_config(rng_mode=_lib.RNG_TAPE, trace_cap=400000, hash_pops=1)
    sim = BatchedSim(cfg, 2)
    sim.reset_tape(*oracle_tapes([o, o]))
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(2):
        assert_env_equals_oracle(sim, e, o, n, st)


def test_rmsc03_with_pov_execution_agent_tape_replay():
    """BASELINE.json configs[2] "rmsc03 ... with POV execution agent": bit-exact vs the oracle (pinned to a recording of the reference with
    its POVExecutionAgent appended); one environment per seed plus a duplicate."""
    from marl_optimal_execution_b200.sim import rmsc03_config
    NS = 10 ** 9
    pv = dict(pov=0.5, quantity=120000, is_buy=1, start_ns=(9 * 3600 + 32 * 60) * NS, end_ns=(9 * 3600 + 43 * 60) * NS, freq_ns=30 * NS, lookback_ns=30 * NS)
    os_ = [OracleSim(3, seed, TRACE_ALL, pov_exec=pv) for seed in (123456789, 1001, 123456789)]
    ns = [o.run() for o in os_]
    cfg = rmsc03_config(pov_exec=True, rng_mode=_lib.RNG_TAPE, trace_cap=400000, hash_pops=1)
    sim = BatchedSim(cfg, 3)
    sim.reset_tape(*oracle_tapes(os_))
    sim.run()
    sim.finalize()
    st = sim.stats()
    assert list(st["messages"]) == ns and ns[0] == 161747 and (st["flags"] == _lib.F_DONE).all(), (st["messages"], st["flags"])
    for e, o in enumerate(os_):
        assert int(st["pop_hash"][e]) == o.pop_hash()
        p, nt, sn = sim.split_trace(e)
        assert np.array_equal(p, o.trace("pops")) and np.array_equal(nt, o.trace("notes")) and np.array_equal(sn, o.trace("snaps"))
        assert np.array_equal(sim.holdings(e)[:, :4], o.holdings()[:, :4]) and np.array_equal(sim.pov_exec(e), o.pov_exec())
    assert (st["sum_shares"] == 0).all() and (st["sum_cash"] == 64 * 10 ** 7).all()


# ---- parametrised parity: the oracle and the CUDA simulator built from the SAME mutated abx_sim_config (tests/param_cases.py) ----
from helpers import assert_env_equals_oracle       # noqa: E402
from param_cases import RMSC03_CASES, SPARSE_ZI_CASES   # noqa: E402


def _run_param_case(cfg, seed, holdings_cols):
    o = OracleSim.from_config(cfg, seed, TRACE_ALL)
    n = o.run()
    sim = BatchedSim(cfg, 3)
    sim.reset_tape(*oracle_tapes([o, o, o]))
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(3):
        assert_env_equals_oracle(sim, e, o, n, st, traces=(e == 2), holdings_cols=holdings_cols)
    return n


@pytest.mark.parametrize("variant,mutate,seed", SPARSE_ZI_CASES, ids=lambda v: getattr(v, "__name__", str(v)))
def test_sparse_zi_non_default_parameters(variant, mutate, seed):
    cfg = sparse_zi_config(variant, rng_mode=_lib.RNG_TAPE, trace_cap=1500000 if variant == 1000 else 200000, hash_pops=1)
    mutate(cfg)
    assert _run_param_case(cfg, seed, 5) > 3000


@pytest.mark.parametrize("pov,mutate,seed", RMSC03_CASES, ids=lambda v: getattr(v, "__name__", str(v)))
def test_rmsc03_non_default_parameters(pov, mutate, seed):
    from marl_optimal_execution_b200.sim import rmsc03_config
    cfg = rmsc03_config(pov_exec=pov, rng_mode=_lib.RNG_TAPE, trace_cap=700000, hash_pops=1)
    mutate(cfg)
    assert _run_param_case(cfg, seed, 4) > 5000
