"""The config-as-script surface: an `abides.py -c <config> <flags>` command line builds the same parameter set the reference's config script would hand
to Kernel.runner (flags of config/rmsc03.py:30-46, config/sparse_zi_1000.py:17-34, config/marketreplay.py:19-28), checked through the CPU emulation
against the oracle built from the mutated struct."""
import numpy as np
import pytest

from helpers import assert_env_equals_oracle, build_emu, oracle_tapes
from marl_optimal_execution_b200 import _lib, configs
from marl_optimal_execution_b200.sim import BatchedSim
from oracle.oracle import OracleSim, TRACE_ALL


@pytest.fixture(scope="module")
def emu():
    return build_emu()


def test_rmsc03_market_maker_flags(emu):
    cfg, run = configs.from_argv(["-c", "rmsc03", "-t", "ABM", "-d", "20190628", "-s", "77", "--mm-pov", "0.1", "--mm-num-ticks", "8", "--mm-window-size", "3",
                                  "--mm-min-order-size", "11", "--mm-wake-up-freq", "2S", "-l", "x"], lib=_lib.load(emu))
    assert (cfg.mm_pov, cfg.mm_num_ticks, cfg.mm_window_size, cfg.mm_min_order_size, cfg.mm_wake_ns, run.seed) == (0.1, 8, 3, 11, 2 * 10 ** 9, 77)
    cfg.rng_mode, cfg.trace_cap, cfg.hash_pops = _lib.RNG_TAPE, 400000, 1
    o = OracleSim.from_config(cfg, run.seed, TRACE_ALL)
    n = o.run()
    sim = BatchedSim(cfg, 1, lib_path=emu)
    sim.reset_tape(*oracle_tapes([o]))
    sim.run(); sim.finalize()
    assert_env_equals_oracle(sim, 0, o, n, sim.stats(), holdings_cols=4)


def test_sparse_zi_flags_and_seeding(emu):
    cfg, run = configs.from_argv(["-c", "sparse_zi_100", "-s", "123456789", "-b", "0", "-l", "golden"], lib=_lib.load(emu))
    sim = run(n_envs=3, lib_path=emu)
    sim.run(); sim.finalize()
    st = sim.stats()
    assert (st["flags"] == _lib.F_DONE).all() and len(set(st["messages"].tolist())) == 3 and (st["sum_shares"] == 0).all()
    with pytest.raises(SystemExit):
        configs.from_argv(["-c", "twoSymbols"])
    assert configs.timedelta_ns("1S") == 10 ** 9 and configs.timedelta_ns("30s") == 30 * 10 ** 9 and configs.timedelta_ns("1min") == 60 * 10 ** 9


def test_rmsc01_and_rmsc02_command_lines():
    """`abides.py -c rmsc01 -s 7` / `-c rmsc02`: the presets of config/rmsc01.py / config/rmsc02.py; an --agent_name (a Python class under test) is rejected."""
    import pytest
    from marl_optimal_execution_b200 import _lib, configs
    from helpers import build_emu
    L = _lib.load(build_emu())
    c1, run1 = configs.from_argv(["-c", "rmsc01", "-s", "7"], lib=L)
    c2, run2 = configs.from_argv(["-c", "rmsc02", "-l", "x"], lib=L)
    assert (c1.population, c1.n_agents, c1.hbl_L, c1.latency_model, c1.mkm_subscribe) == (3, 101, 2, _lib.LAT_ZERO if hasattr(_lib, "LAT_ZERO") else 2, 0) and run1.seed == 7
    assert (c2.mkm_subscribe, c2.mom_subscribe, c2.n_noise, c2.start_ns, c2.stop_ns) == (1, 1, 6, 0, 17 * 3600 * 10 ** 9) and run2.seed == 0
    with pytest.raises(SystemExit):
        configs.from_argv(["-c", "rmsc01", "-a", "MyAgent"], lib=L)
