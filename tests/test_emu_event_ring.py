"""The exchange event ring (cfg.event_ring_cap: order arrivals + the BEST_BID / BEST_ASK / LAST_TRADE log lines of util/OrderBook.py:114-141) and its torch
reductions, on the CPU emulation of the product logic: the ring must tell exactly what the oracle's traces tell."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import build_emu, oracle_tapes
from marl_optimal_execution_b200 import _lib, realism as R
from marl_optimal_execution_b200.sim import BatchedSim, rmsc03_config, sparse_zi_config
from oracle.oracle import OracleSim, TRACE_ALL

EXEC = 8


@pytest.fixture(scope="module")
def emu():
    return build_emu()


def _ring(sim, L, cap):
    raw = torch.zeros(sim.n_envs, cap, 4, dtype=torch.int32)
    cnt = torch.zeros(sim.n_envs, dtype=torch.int32)
    _lib.check(L, L.abx_sim_events_device(sim._h, C.c_void_p(raw.data_ptr()), C.c_void_p(cnt.data_ptr()), None), "abx_sim_events_device")
    return R.unroll_events(raw, cnt)


def test_ring_equals_oracle_traces_rmsc03(emu):
    """rmsc03 (variable order sizes): order arrivals == the oracle's book operations, LAST_TRADE quantities == the fills of each incoming order, BEST_BID /
    BEST_ASK == the book snapshot after each limit order."""
    L = _lib.load(emu)
    o = OracleSim(3, 1001, TRACE_ALL)
    o.run()
    cap = 1 << 18
    sim = BatchedSim(rmsc03_config(lib=L, rng_mode=_lib.RNG_TAPE, hash_pops=1, event_ring_cap=cap), 1, lib_path=emu)
    sim.reset_tape(*oracle_tapes([o]))
    sim.run(); sim.finalize()
    t, kind, a, b, valid = (x[0] for x in _ring(sim, L, cap))
    assert int(valid.sum()) < cap
    ops, notes, snaps = o.trace("ops"), o.trace("notes"), o.trace("snaps")
    lim = ops[ops[:, 1] == 0]
    od = valid & (kind == R.EV_ORDER)
    assert int(od.sum()) == len(lim) == o.counter("limit")
    assert np.array_equal(t[od].numpy(), lim[:, 0]) and np.array_equal(a[od].numpy(), lim[:, 5]) and np.array_equal(b[od].numpy(), np.where(lim[:, 4] == 1, lim[:, 6], -lim[:, 6]))
    ex = notes[notes[:, 2] == EXEC]
    tr = valid & (kind == R.EV_LAST_TRADE)
    assert int(b[tr].sum()) == int(ex[:, 5].sum()) // 2 and int(tr.sum()) > 100          # every fill is reported to both sides
    assert int(a[tr][-1]) == o.book_l1()[4]
    # BEST_BID / BEST_ASK after each handleLimitOrder == the recorded book snapshot rows of those operations
    lim_snaps = snaps[ops[:, 1] == 0]
    bb = valid & (kind == R.EV_BEST_BID)
    assert np.array_equal(a[bb].numpy(), lim_snaps[lim_snaps[:, 0] > 0][:, 3]) and np.array_equal(b[bb].numpy(), lim_snaps[lim_snaps[:, 0] > 0][:, 4])
    ba = valid & (kind == R.EV_BEST_ASK)
    assert np.array_equal(a[ba].numpy(), lim_snaps[lim_snaps[:, 1] > 0][:, 9]) and np.array_equal(b[ba].numpy(), lim_snaps[lim_snaps[:, 1] > 0][:, 10])


def test_bars_and_order_flow_from_the_ring(emu):
    """sparse_zi_100 under Philox: bars reduced from the ring == bars of the minute-by-minute stepped simulation; order-flow counts == the counters."""
    L = _lib.load(emu)
    cap = 16384
    cfg = sparse_zi_config(100, lib=L, event_ring_cap=cap)
    sim = BatchedSim(cfg, 3, lib_path=emu)
    sim.reset([5, 6, 7])
    close_s, vol_s = R.minute_bars(sim, 390)
    sim.run(); sim.finalize()
    st = sim.stats()
    t, kind, a, b, valid = _ring(sim, L, cap)
    close, vol = R.bars_from_events(t, kind, a, b, valid, cfg.mkt_open_ns, 390, open_price=cfg.r_bar)
    assert np.array_equal(close.numpy(), close_s) and np.array_equal(vol.numpy(), vol_s)
    f = R.order_flow_facts(t, kind, valid, cfg.mkt_open_ns, cfg.mkt_close_ns, binwidth_s=60)
    assert np.array_equal(f["n_orders"].numpy(), st["limit_orders"].astype(np.int64))
    assert f["bin_counts"].shape == (3, 390) and (f["bin_counts"].sum(dim=1) <= f["n_orders"]).all() and (f["interarrival_mean"] > 1.0).all()
    assert (f["interarrival_log10_hist"].sum(dim=1) == f["n_orders"] - 1).all()
    sp = R.spread_facts(t, kind, a, valid)
    assert (sp["spread_mean"] > 0).all() and np.array_equal(sp["spread_last"].numpy(), (st["best_ask"] - st["best_bid"]).astype(np.int64))
    m = R.all_metrics(close.numpy(), vol.numpy())
    assert m["returns"].shape == (3, 389) and np.isfinite(m["kurtosis"]).all()


def test_ring_wraps(emu):
    L = _lib.load(emu)
    cap = 1024
    cfg = sparse_zi_config(100, lib=L, event_ring_cap=cap)
    sim = BatchedSim(cfg, 1, lib_path=emu)
    sim.reset([9])
    sim.run()
    t, kind, a, b, valid = (x[0] for x in _ring(sim, L, cap))
    assert bool(valid.all()) and bool((t[1:] >= t[:-1]).all())                          # the last `cap` events, oldest first
    assert int(a[kind == R.EV_LAST_TRADE][-1]) == int(sim.stats()["last_trade"][0])
