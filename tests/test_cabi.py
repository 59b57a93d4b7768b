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
