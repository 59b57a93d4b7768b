"""The Philox-mode logic (what the benchmarked kernels run) against the oracle, on the CPU emulation of the product's warp-uniform code
(tests/emu, a test tool): every standard variate an environment draws is logged (cfg.draw_log_cap), the oracle is re-run on exactly those
draws (external tapes) from the same abx_sim_config, and pops, exchange messages, book snapshots, counters and holdings must be equal.
The GPU suite runs the same check on the CUDA kernels (tests/test_gpu_philox_oracle.py)."""
import numpy as np
import pytest

from helpers import assert_env_equals_oracle, build_emu, oracle_rerun_of_philox_env
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.sim import BatchedSim, rmsc01_config, rmsc02_config, rmsc03_config, sparse_zi_config
from oracle.oracle import TRACE_ALL


@pytest.fixture(scope="module")
def emu():
    return build_emu()


@pytest.mark.parametrize("variant,seeds", [(100, [5, 6, 77]), (1000, [123])])
def test_sparse_zi_philox_run_equals_oracle_on_its_own_draws(emu, variant, seeds):
    big = variant == 1000
    cfg = sparse_zi_config(variant, lib=_lib.load(emu), trace_cap=700000 if big else 70000, hash_pops=1, draw_log_cap=500000 if big else 60000)
    sim = BatchedSim(cfg, len(seeds), lib_path=emu)
    sim.reset(seeds)
    init = [sim.agent_init(e) for e in range(len(seeds))]
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(len(seeds)):
        o, n = oracle_rerun_of_philox_env(sim, e, init[e], TRACE_ALL)
        assert_env_equals_oracle(sim, e, o, n, st)
    assert len(set(int(x) for x in st["pop_hash"])) == len(seeds)        # different seeds, different days


@pytest.mark.parametrize("pov", [False, True])
def test_rmsc03_philox_run_equals_oracle_on_its_own_draws(emu, pov):
    cfg = rmsc03_config(lib=_lib.load(emu), pov_exec=pov, trace_cap=500000, hash_pops=1, draw_log_cap=60000)
    sim = BatchedSim(cfg, 2, lib_path=emu)
    sim.reset([11, 12])
    init = [sim.agent_init(e) for e in range(2)]
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(2):
        o, n = oracle_rerun_of_philox_env(sim, e, init[e], TRACE_ALL)
        if int(st["flags"][e]) & _lib.F_OBS_INVALID:      # the POV agent met an empty book side: the reference raises there, nothing to compare
            continue
        assert_env_equals_oracle(sim, e, o, n, st, holdings_cols=4)


def test_rmsc01_philox_run_equals_oracle_on_its_own_draws(emu):
    """config/rmsc01.py population seeded by Philox: the oracle, re-run on the environment's own draws, sees the same HBL limit prices, market-maker
    ladders and trades (09:30 - 09:36)."""
    stop = (9 * 3600 + 36 * 60) * 10 ** 9
    cfg = rmsc01_config(lib=_lib.load(emu), trace_cap=300000, hash_pops=1, draw_log_cap=60000, stop_ns=stop)
    sim = BatchedSim(cfg, 2, lib_path=emu)
    sim.reset([31, 32])
    init = [sim.agent_init(e) for e in range(2)]
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(2):
        o, n = oracle_rerun_of_philox_env(sim, e, init[e], TRACE_ALL)
        assert_env_equals_oracle(sim, e, o, n, st)
    assert st["pop_hash"][0] != st["pop_hash"][1] and st["fills"].min() > 5              # rmsc01 trades rarely: ~40 fills in the reference's first 15 minutes


def test_rmsc02_philox_run_equals_oracle_on_its_own_draws(emu):
    """config/rmsc02.py seeded by Philox (latency matrix drawn on the device, kernel-noise stream, subscriptions): oracle re-run on the logged draws, midnight - 11:00."""
    stop = 11 * 3600 * 10 ** 9
    cfg = rmsc02_config(lib=_lib.load(emu), trace_cap=300000, hash_pops=1, draw_log_cap=100000, stop_ns=stop)
    sim = BatchedSim(cfg, 2, lib_path=emu)
    sim.reset([41, 42])
    init = [sim.agent_init(e) for e in range(2)]
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(2):
        o, n = oracle_rerun_of_philox_env(sim, e, init[e], TRACE_ALL)
        assert_env_equals_oracle(sim, e, o, n, st)
    assert st["pop_hash"][0] != st["pop_hash"][1] and st["fills"].min() > 50
