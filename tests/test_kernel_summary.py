"""Kernel.runner's stdout contract (Kernel.py:321-343, agent/TradingAgent.py:115-138) from the host mirror:
"Final holdings for ...", "Event Queue elapsed: ..., messages: N, messages per second: R", "Mean ending value by agent type"."""
import os
import re

import numpy as np
import pytest

from helpers import build_emu, oracle_tapes
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.sim import BatchedSim, agent_directory, format_kernel_summary, rmsc03_config, sparse_zi_config
from oracle.oracle import OracleSim, TRACE_ALL

REF_TXT = "/root/reference/tests/sparse_zi_1000.txt"


@pytest.fixture(scope="module")
def emu():
    return build_emu()


@pytest.mark.skipif(not os.path.exists(REF_TXT), reason="reference tree not present")
def test_summary_text_equals_the_references_recorded_stdout(emu):
    """Every 'Final holdings' line and every 'Mean ending value' line of the reference's own capture of sparse_zi_1000 (seed 123456789),
    character for character, from the pinned oracle's final holdings."""
    txt = open(REF_TXT).read()
    s = OracleSim(1000, 123456789, 0)
    n = s.run()
    cfg = sparse_zi_config(1000, lib=_lib.load(emu))
    names, types = agent_directory(cfg)
    assert len(names) == 1000
    lines = format_kernel_summary(names, types, s.holdings(), n, int(cfg.starting_cash), "JPM", elapsed_s=59.733625)
    ref_lines = [ln.strip() for ln in txt.splitlines()]
    ref_final = [ln for ln in ref_lines if ln.startswith("Final holdings for")]
    ours_final = [ln for ln in lines if ln.startswith("Final holdings for")]
    assert len(ref_final) == 1000 and sorted(ours_final) == sorted(ref_final)
    ref_mean = [ln for ln in ref_lines if re.match(r"ZeroIntelligenceAgent Type \d .*: -?\d+$", ln)]
    i = lines.index("Mean ending value by agent type:")
    assert len(ref_mean) == 7 and lines[i + 1: i + 8] == ref_mean and lines[-1] == "Simulation ending!"
    assert lines[i - 1] == "Event Queue elapsed: 0 days 00:00:59.733625, messages: 185200, messages per second: 3100.4"
    assert lines[i - 1] in ref_lines


def test_summary_through_the_product_path_matches_oracle(emu):
    """BatchedSim.kernel_summary (abx_sim_stats + abx_sim_holdings) on a replayed sparse_zi_100 day == the same text from the oracle's run."""
    o = OracleSim(100, 1001, TRACE_ALL)                   # the RNG tapes are part of the trace
    n = o.run()
    cfg = sparse_zi_config(100, lib=_lib.load(emu), rng_mode=_lib.RNG_TAPE)
    sim = BatchedSim(cfg, 1, lib_path=emu)
    sim.reset_tape(*oracle_tapes([o]))
    sim.run()
    sim.finalize()
    assert int(sim.stats()[0]["flags"]) == _lib.F_DONE
    names, types = agent_directory(cfg)
    want = format_kernel_summary(names, types, o.holdings(), n, int(cfg.starting_cash), "JPM", elapsed_s=2.0)
    got = sim.kernel_summary(0, "JPM", elapsed_s=2.0)
    assert got == want and len(got) == 100 + 1 + 1 + 7 + 1
    assert got[100] == "Event Queue elapsed: 0 days 00:00:02, messages: %d, messages per second: %.1f" % (n, n / 2.0)


def test_rmsc03_directory_follows_the_config_script(emu):
    L = _lib.load(emu)
    names, types = agent_directory(rmsc03_config(L, pov_exec=True))
    assert len(names) == 64 and names[0] == "NoiseAgent 1" and names[49] == "NoiseAgent 50" and names[50] == "Value Agent 51"
    assert names[60] == "POV_MARKET_MAKER_AGENT_61" and names[61:63] == ["MOMENTUM_AGENT_62", "MOMENTUM_AGENT_63"] and names[63] == "POV_EXECUTION_AGENT"
    assert [t for i, t in enumerate(types) if i == 0 or types[i - 1] != t] == ["NoiseAgent", "ValueAgent", "POVMarketMakerAgent", "MomentumAgent", "ExecutionAgent"]
    rows = np.array([[i + 1, 0, 10000000 + i, 10000000 + i, 0] for i in range(64)])
    lines = format_kernel_summary(names, types, rows, 5, 10000000)
    assert lines[0] == "Final holdings for NoiseAgent 1: { CASH: 10000000 }.  Marked to market: 10000000"
    assert lines[64:] == ["Mean ending value by agent type:", "NoiseAgent: 24", "ValueAgent: 54", "POVMarketMakerAgent: 60", "MomentumAgent: 62",
                          "ExecutionAgent: 63", "Simulation ending!"]
