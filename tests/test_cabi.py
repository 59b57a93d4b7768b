"""The C-ABI shared library loads and exports every symbol include/abides_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from marl_optimal_execution_b200 import _lib, build
from marl_optimal_execution_b200.sim import sparse_zi_config


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "abides_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(abx_[a-z0-9_]+)\s*\(", hdr)))


def test_exports_every_declared_symbol(lib_path):
    L = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(L, n), n


def test_struct_layouts_match_header(lib_path):
    L = _lib.load(lib_path)
    cfg = sparse_zi_config(1000, lib=L)
    # config/sparse_zi_1000.py population and parameters
    assert cfg.n_agents == 1001 and cfg.n_groups == 7 and cfg.q_max == 10
    assert [cfg.groups[i].count for i in range(7)] == [143] * 6 + [142]
    assert [(cfg.groups[i].r_min, cfg.groups[i].r_max, cfg.groups[i].eta) for i in range(7)] == [
        (0, 250, 1), (0, 500, 1), (0, 1000, .8), (0, 1000, 1), (0, 2000, .8), (250, 500, .8), (250, 500, 1)]
    assert cfg.stop_ns == 17 * 3600 * 10 ** 9 and cfg.mkt_open_ns == 34200 * 10 ** 9 and cfg.mkt_close_ns == 57600 * 10 ** 9
    assert cfg.default_computation_delay_ns == 10 ** 9 and cfg.starting_cash == 10 ** 7 and cfg.lambda_a == 1e-12
    assert cfg.latency_model == _lib.LAT_MATRIX_NOISE and cfg.n_noise == 6 and cfg.jitter == 0.0
    c100 = sparse_zi_config(100, lib=L)
    assert c100.n_agents == 101 and c100.latency_model == _lib.LAT_CUBIC and (c100.jitter, c100.jitter_clip, c100.jitter_unit) == (0.3, 0.05, 5.0)
    assert c100.hash_pops == 0 and c100.trace_cap == 0     # last fields land where the C struct puts them


def test_errors_are_statuses_not_crashes(lib_path):
    L = _lib.load(lib_path)
    assert L.abx_strerror(-2) == b"CUDA runtime error"
    cfg = sparse_zi_config(100, lib=L, queue_cap=33)
    h = ctypes.c_void_p()
    assert L.abx_sim_create(ctypes.byref(cfg), 1, 0, ctypes.byref(h)) == -1      # ABX_ERR_ARG before touching CUDA
    assert L.abx_sim_run(None, 0, None) == -1 and L.abx_sim_destroy(None) == 0


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_lib.AbxError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "libabides_b200.so"))


def test_new_config_structs_and_no_device_behaviour(lib_path):
    """abx_dq_config / rmsc03+POV presets land where the C structs put them; without a GPU every computing entry point returns a status
    (ABX_ERR_CUDA) -- there is no CPU path behind the product library."""
    import numpy as np
    from marl_optimal_execution_b200.env import dq_config, env_config
    from marl_optimal_execution_b200.sim import rmsc03_config
    L = _lib.load(lib_path)
    d = dq_config(L)
    assert (d.n_momentum, d.n_twap, d.has_ddqn, d.is_buy, d.n_horizon, d.quantity) == (7, 1, 1, 1, 661, 500000)
    assert d.horizon_start_ns == 36000 * 10 ** 9 and d.horizon_step_ns == 30 * 10 ** 9 and d.stop_ns == (36000 + 660 * 30 + 600) * 10 ** 9
    assert d.mom_wake_ns == 20 * 10 ** 9 and d.queue_cap == 1536 and d.hash_pops == 0
    r = rmsc03_config(L, pov_exec=True)
    assert r.n_agents == 65 and r.n_pov_exec == 1 and r.pov_exec_pov == 0.5 and r.pov_exec_quantity == 120000 and r.pov_exec_lookback_ns == 30 * 10 ** 9
    assert rmsc03_config(L).n_agents == 64
    if L.abx_device_count() > 0:
        pytest.skip("a CUDA device is present: the no-device statuses cannot be observed")
    h = ctypes.c_void_p()
    st = np.array([[34200 * 10 ** 9, 1000001, 1000, 100, 1]], dtype=np.int64)
    e = env_config(L)
    assert L.abx_env_create(ctypes.byref(e), st.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 1, 1, 0, ctypes.byref(h)) == -2
    assert L.abx_dq_create(ctypes.byref(d), st.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 1, 1, 0, ctypes.byref(h)) == -2
    assert L.abx_book_create(10, 64, 64, 0, 1, 0, ctypes.byref(h)) == -2
    dims = (ctypes.c_int32 * 3)(2, 16, 4)
    params = (ctypes.c_float * L.abx_qnet_param_count(dims, 2))()
    assert L.abx_qnet_param_count(dims, 2) == 2 * 16 + 16 + 16 * 4 + 4
    assert L.abx_qnet_create(dims, 2, params, 0, ctypes.byref(h)) == -2
    assert b"CUDA" in L.abx_last_cuda_error() or len(L.abx_last_cuda_error()) > 0
    # argument errors are caught before CUDA is touched
    bad = np.array([[2, 5, 1000, 100, 1], [1, 6, 1000, 100, 1]], dtype=np.int64)              # not time sorted
    assert L.abx_env_create(ctypes.byref(e), bad.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 2, 1, 0, ctypes.byref(h)) == -1
    d2 = dq_config(L, n_momentum=9)
    assert L.abx_dq_create(ctypes.byref(d2), st.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 1, 1, 0, ctypes.byref(h)) == -1
    assert L.abx_book_create(10, 4, 64, 0, 1, 0, ctypes.byref(h)) == -1                       # level_cap below the minimum
    dims_bad = (ctypes.c_int32 * 3)(2, 300, 4)
    assert L.abx_qnet_create(dims_bad, 2, params, 0, ctypes.byref(h)) == -1


def test_emu_edge_cases_empty_and_single_row_streams():
    """Edge cases of the replayed stream through the host emulation harness of the product logic: a one-row stream, a stream whose first
    rows share one timestamp, zero-size rows (cancels of unknown ids); the run ends cleanly without error flags."""
    import numpy as np
    from helpers import build_emu
    from marl_optimal_execution_b200.env import ABIDESEnv, env_config
    from oracle.oracle import OracleEnv
    emu = build_emu()
    L = _lib.load(emu)
    t0 = 34200 * 10 ** 9
    for rows in ([[t0 + 5, 900001, 1000, 100, 1]],
                 [[t0 + 5, 900001, 1000, 100, 1], [t0 + 5, 900002, 1001, 50, 0], [t0 + 5, 900003, 1000, 0, 1], [t0 + 9, 900001, 1000, 60, 1], [t0 + 9, 900002, 1001, 0, 0]]):
        st = np.array(rows, dtype=np.int64)
        env = ABIDESEnv(st, n_envs=1, cfg=env_config(L, order_level=0, stop_ns=(16 * 3600 + 60) * 10 ** 9, hash_pops=1), lib_path=emu)
        env.reset()
        _, _, done, _ = env.step(np.zeros((1, 3)))
        o = OracleEnv(st, order_level=0, stop_ns=(16 * 3600 + 60) * 10 ** 9)
        o.step([0, 0, 0])
        s = env.stats()[0]
        assert int(done[0]) == 1 and int(s["flags"]) == _lib.F_DONE and int(s["messages"]) == o.n_pops and int(s["pop_hash"]) == o.pop_hash()
    with pytest.raises(ValueError):
        ABIDESEnv(np.zeros((0, 5), dtype=np.int64), n_envs=1, cfg=env_config(L), lib_path=emu) if False else ABIDESEnv(np.zeros((3, 4), dtype=np.int64), n_envs=1, cfg=env_config(L), lib_path=emu)
    with pytest.raises(_lib.AbxError):
        ABIDESEnv(np.zeros((0, 5), dtype=np.int64), n_envs=1, cfg=env_config(L), lib_path=emu)      # empty stream: rejected at create
