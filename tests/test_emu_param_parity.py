"""Parametrised parity (CPU suite: the product's warp-uniform logic compiled by tests/emu): non-default agent counts, arrival rates, oracle
parameters, q_max, order sizes, latency-model parameters, market-maker settings -- the oracle and the product are built from the SAME mutated
abx_sim_config and must agree bit for bit (pops, exchange messages, snapshots, holdings).  The GPU suite repeats the cases on the CUDA kernels."""
import numpy as np
import pytest

from helpers import assert_env_equals_oracle, build_emu, oracle_tapes
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.sim import BatchedSim, rmsc03_config, sparse_zi_config
from oracle.oracle import OracleSim, TRACE_ALL
from param_cases import RMSC03_CASES, SPARSE_ZI_CASES


@pytest.fixture(scope="module")
def emu():
    return build_emu()


def run_case(cfg, seed, lib_path, holdings_cols):
    o = OracleSim.from_config(cfg, seed, TRACE_ALL)
    n = o.run()
    sim = BatchedSim(cfg, 2, lib_path=lib_path)
    sim.reset_tape(*oracle_tapes([o, o]))
    sim.run()
    sim.finalize()
    st = sim.stats()
    for e in range(2):
        assert_env_equals_oracle(sim, e, o, n, st, traces=(e == 1), holdings_cols=holdings_cols)
    return n, o


@pytest.mark.parametrize("variant,mutate,seed", SPARSE_ZI_CASES, ids=lambda v: getattr(v, "__name__", str(v)))
def test_sparse_zi_parameters(emu, variant, mutate, seed):
    cfg = sparse_zi_config(variant, lib=_lib.load(emu), rng_mode=_lib.RNG_TAPE, trace_cap=1500000 if variant == 1000 else 200000, hash_pops=1)
    mutate(cfg)
    n, o = run_case(cfg, seed, emu, 5)
    assert n > 3000 and o.counter("fills") > 0


@pytest.mark.parametrize("pov,mutate,seed", RMSC03_CASES, ids=lambda v: getattr(v, "__name__", str(v)))
def test_rmsc03_parameters(emu, pov, mutate, seed):
    cfg = rmsc03_config(lib=_lib.load(emu), pov_exec=pov, rng_mode=_lib.RNG_TAPE, trace_cap=700000, hash_pops=1)
    mutate(cfg)
    n, o = run_case(cfg, seed, emu, 4)
    assert n > 5000 and o.counter("limit") > 1000


def test_default_struct_equals_the_config_scripts():
    """The oracle's own statement of the config scripts' numbers == the product's presets (field by field, capacities aside)."""
    import ctypes as C
    from oracle.oracle import lib
    skip = {"queue_cap", "level_cap", "order_cap", "rng_mode", "trace_cap", "hash_pops", "draw_log_cap", "event_ring_cap", "_pad", "_pad0", "_pad1", "groups"}
    for variant, mk in ((100, lambda: sparse_zi_config(100)), (1000, lambda: sparse_zi_config(1000)), (3, lambda: rmsc03_config()), (4, lambda: rmsc03_config(pov_exec=True))):
        a, b = mk(), _lib.SimConfig()
        assert lib().abo_default_config(variant, C.addressof(b)) == 0
        for name, _ in _lib.SimConfig._fields_:
            if name not in skip:
                assert getattr(a, name) == getattr(b, name), (variant, name)
        for g in range(8):
            assert tuple(getattr(a.groups[g], f) for f in ("count", "r_min", "r_max", "eta")) == tuple(getattr(b.groups[g], f) for f in ("count", "r_min", "r_max", "eta"))
