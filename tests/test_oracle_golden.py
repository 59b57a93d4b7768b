"""Pin the oracle (oracle/abides_oracle.c) against traces recorded from the live reference
(tools/record_reference.py -> tests/golden/*.npz) and against the reference's own recorded run
tests/sparse_zi_1000.txt."""
import os
import re

import numpy as np
import pytest

from oracle.oracle import OracleSim, TRACE_ALL


def test_z100_full_trace_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "z100_s123456789_full.npz"))
    s = OracleSim(100, 123456789, TRACE_ALL)
    assert s.run() == int(g["n_pops"]) == 18871
    for k in ("pops", "ops", "notes", "snaps"):
        assert np.array_equal(s.trace(k), g[k]), k
    assert np.array_equal(s.holdings(), g["holdings"])
    off = g["tape_offsets"]
    assert s.n_streams == len(off) - 1
    for i in range(s.n_streams):
        kinds, bits = s.tape(i)
        assert s.stream_seed(i) == g["stream_seeds"][i]
        assert np.array_equal(kinds, g["tape_kind"][off[i]:off[i + 1]].view(np.uint8)), i
        assert np.array_equal(bits, g["tape_bits"][off[i]:off[i + 1]]), i
    assert np.array_equal(s.global_exp_tape(), g["global_exp_tape"])


@pytest.mark.parametrize("variant,seed,name", [(1000, 123456789, "z1000_s123456789"), (1000, 1001, "z1000_s1001"),
                                               (100, 1001, "z100_s1001")])
def test_digest_bit_exact(golden_dir, variant, seed, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    s = OracleSim(variant, seed, 0)
    assert s.run() == int(g["n_pops"])
    assert np.array_equal(s.hash_ckpt(), g["pop_hash_ckpt"])           # event order, every 1000 pops
    assert s.note_hash() == int(g["note_hash"])                        # every outbound exchange message (fills, L1 replies)
    assert s.snap_hash() == int(g["snap_hash"])                        # book state after every book op
    assert np.array_equal(s.holdings(), g["holdings"])
    assert s.counter("max_bid_levels") == g["max_levels"][0] and s.counter("max_ask_levels") == g["max_levels"][1]
    assert s.counter("max_resting") == int(g["max_resting"])


def test_z1000_known_answers():
    s = OracleSim(1000, 123456789, 0)
    assert s.run() == 185200                                           # tests/sparse_zi_1000.txt:22
    # SURVEY App. B.3 shape facts
    assert s.counter("limit") == 24416 and s.counter("fills") == 5459 and s.counter("cancel") == 12975
    assert s.counter("spread_queries") == 25424 and s.counter("max_queue") == 2004


@pytest.mark.skipif(not os.path.exists("/root/reference/tests/sparse_zi_1000.txt"), reason="reference tree not present")
def test_z1000_matches_reference_recorded_stdout():
    """The reference's own golden file: 1000 'Final holdings' lines + message count."""
    txt = open("/root/reference/tests/sparse_zi_1000.txt").read()
    assert re.search(r"messages: (\d+)", txt).group(1) == "185200"
    rows = re.findall(r"Final holdings for ZI Agent (\d+) .*?\{ (?:JPM: (-?\d+), )?CASH: (-?\d+) \}\.  Marked to market: (-?\d+)", txt)
    assert len(rows) == 1000
    ref = {int(a): (int(sh or 0), int(c), int(m)) for a, sh, c, m in rows}
    s = OracleSim(1000, 123456789, 0)
    s.run()
    for aid, sh, cash, mtm, _ in s.holdings():
        assert ref[int(aid)] == (sh, cash, mtm), aid


@pytest.mark.parametrize("seed", [123456789, 1001])
def test_rmsc03_digest_bit_exact(golden_dir, seed):
    """config/rmsc03.py (50 Noise + 10 Value + POV market maker + 2 Momentum, with the one-line getTransactedVolume alias the
    shipped config needs): event order, exchange messages incl. transacted-volume driven ladder sizes, book snapshots,
    every RNG draw incl. the GLOBAL np.random stream, final holdings."""
    path = os.path.join(golden_dir, "rmsc03_s%d.npz" % seed)
    if not os.path.exists(path):
        pytest.skip("fixture not recorded")
    g = np.load(path)
    s = OracleSim(3, seed, 16)
    assert s.run() == int(g["n_pops"])
    assert np.array_equal(s.hash_ckpt(), g["pop_hash_ckpt"])
    assert s.note_hash() == int(g["note_hash"]) and s.snap_hash() == int(g["snap_hash"])
    assert np.array_equal(s.holdings()[:, :4], g["holdings"][:, :4])
    assert [s.stream_seed(i) for i in range(s.n_streams)] == list(g["stream_seeds"])
    assert [len(s.tape(i)[0]) for i in range(s.n_streams)] == list(g["stream_draws"])
    gk, gb = s.global_tape()
    assert np.array_equal(gk, g["global_kind"].view(np.uint8)) and np.array_equal(gb, g["global_bits"])
    assert s.counter("max_bid_levels") == g["max_levels"][0] and s.counter("max_resting") == int(g["max_resting"])


def test_rmsc03_with_pov_execution_agent(golden_dir):
    """config/rmsc03.py with the reference's POVExecutionAgent appended by the recorder (tools/record_reference.py --pov-exec): the
    "rmsc03 ... with POV execution agent" shape of BASELINE.json configs[2]."""
    g = np.load(os.path.join(golden_dir, "rmsc03_pov_s123456789.npz"))
    NS = 10 ** 9
    pv = dict(pov=float(g["pov_exec"][0]), quantity=int(g["pov_exec"][1]), is_buy=int(g["pov_exec"][2]), start_ns=(9 * 3600 + 32 * 60) * NS,
              end_ns=(9 * 3600 + 43 * 60) * NS, freq_ns=30 * NS, lookback_ns=30 * NS)
    s = OracleSim(3, 123456789, 31, pov_exec=pv)
    assert s.run() == int(g["n_pops"]) == 161747
    assert np.array_equal(s.hash_ckpt(), g["pop_hash_ckpt"])
    assert s.note_hash() == int(g["note_hash"]) and s.snap_hash() == int(g["snap_hash"])
    assert np.array_equal(s.holdings()[:, :4], g["holdings"][:, :4])
    ops = s.trace("ops")
    assert np.array_equal(ops[ops[:, 2] == 64], g["pov_ops"]) and len(g["pov_ops"]) == 183
    assert list(s.pov_exec()) == [int(g["pov_exec"][3]), int(g["pov_exec"][4]), int(g["pov_exec"][5])]
    gk, gb = s.global_tape()
    assert np.array_equal(gk, g["global_kind"].view(np.uint8)) and np.array_equal(gb, g["global_bits"])


from helpers import oracle_rmsc01_config as rmsc01_config  # noqa: E402


EXEC_AGENT_FIXTURES = [("rmsc03_aggressive_s123456789.npz", 123456789, 67901), ("rmsc03_passive_s123456789.npz", 123456789, 161383), ("rmsc03_passive_limit_s1001.npz", 1001, 161250)]


def exec_agent_config(g):
    """abx_sim_config of config/rmsc03.py + the appended PassiveAgent / AggressiveAgent of a recording (tools/record_reference.py --exec-agent)."""
    import ctypes as C
    from marl_optimal_execution_b200 import _lib
    from oracle.oracle import lib
    xa = g["exec_agent"]
    cfg = _lib.SimConfig()
    assert lib().abo_default_config(4, C.addressof(cfg)) == 0
    cfg.exec_kind, cfg.pov_exec_start_ns, cfg.pov_exec_quantity, cfg.pov_exec_is_buy, cfg.exec_limit_price = int(xa[0]), int(xa[1]), int(xa[2]), int(xa[3]), int(xa[4])
    return cfg


@pytest.mark.parametrize("fixture,seed,pops", EXEC_AGENT_FIXTURES)
def test_rmsc03_with_passive_or_aggressive_agent(golden_dir, fixture, seed, pops):
    """config/rmsc03.py with the reference's AggressiveAgent (BUY 3 000 at 09:36: getCurrentSpread(depth=100) + a market order walked over 23 ask levels, after which
    the POV market maker meets a one-sided book and stops quoting) or PassiveAgent (SELL 800 at the best ask of its own QUERY_SPREAD; BUY 600 at a fixed limit price)
    appended by the recorder: pops, exchange messages, snapshots, the agent's own book operations and all holdings."""
    g = np.load(os.path.join(golden_dir, fixture))
    s = OracleSim.from_config(exec_agent_config(g), seed, TRACE_ALL)
    assert s.run() == int(g["n_pops"]) == pops
    assert np.array_equal(s.hash_ckpt(), g["pop_hash_ckpt"]) and s.note_hash() == int(g["note_hash"]) and s.snap_hash() == int(g["snap_hash"])
    ops = s.trace("ops")
    assert np.array_equal(ops[ops[:, 2] == 64], g["exec_ops"]) and len(g["exec_ops"]) == (23 if int(g["exec_agent"][0]) == 2 else 1)
    assert np.array_equal(s.holdings()[:, :4], g["holdings"][:, :4])


def test_rmsc01_full_trace_bit_exact(golden_dir):
    """config/rmsc01.py run live to 09:45:00 (tools/record_reference.py rmsc01 123456789 --full --stop 09:45:00): every kernel pop,
    every order-book operation incl. the HBL agents' QUERY_ORDER_STREAM-driven limit prices and the MarketMakerAgent's ladder,
    every agent notification, every book snapshot, every stream's draw count and the final holdings."""
    g = np.load(os.path.join(golden_dir, "rmsc01_s123456789_0945.npz"))
    s = OracleSim.from_config(rmsc01_config(), 123456789, TRACE_ALL)
    assert s.run() == int(g["n_pops"]) == 77119
    for name in ("pops", "ops", "notes", "snaps"):
        assert np.array_equal(s.trace(name), g[name]), name
    assert np.array_equal(s.hash_ckpt(), g["pop_hash_ckpt"])
    assert np.array_equal(s.holdings()[:, :4], g["holdings"][:, :4])
    mine, ref = [len(s.tape(i)[0]) for i in range(s.n_streams)], [int(x) for x in g["stream_draws"]]
    assert mine[3:] == ref[3:] and sorted(mine[:3]) == sorted(ref[:3])     # config/rmsc01.py creates its first three streams in another order


def test_rmsc02_full_day_bit_exact(golden_dir):
    """config/rmsc02.py, the WHOLE day run live (tools/record_reference.py rmsc02 123456789 --full): the rmsc01 population with MARKET_DATA subscriptions
    (agent/ExchangeAgent.py:342-387: 25 subscribers served in subscription order after every book operation, at most every 10 s each), the market maker's
    subscription-mode ladder (1-4 levels, levels_quote_dict split), pairwise latency + noise.  117 238 pops, 31 938 book operations -- incl. the quantity a
    CANCEL_ORDER shows at the exchange when the agent booked a partial fill while the message was in flight (the message references the agent's order object)
    -- 80 336 exchange messages with every level of every MARKET_DATA body, every book snapshot, the final holdings."""
    from helpers import oracle_rmsc02_config
    g = np.load(os.path.join(golden_dir, "rmsc02_s123456789.npz"))
    s = OracleSim.from_config(oracle_rmsc02_config(), 123456789, TRACE_ALL)
    assert s.run() == int(g["n_pops"]) == 117238
    for name in ("pops", "ops", "notes", "snaps"):
        assert np.array_equal(s.trace(name), g[name]), name
    assert np.array_equal(s.hash_ckpt(), g["pop_hash_ckpt"])
    assert np.array_equal(s.holdings()[:, :4], g["holdings"][:, :4])
    notes = s.trace("notes")
    assert (notes[:, 2] == 15).sum() > 20000            # MARKET_DATA
