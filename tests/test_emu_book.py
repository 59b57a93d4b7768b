"""CPU CI of the book surface through the host emulation harness of the product logic (a test tool, never a fallback)."""
import pytest

import book_cases
from helpers import build_emu


@pytest.fixture(scope="module")
def emu():
    return build_emu()


def test_kat1(emu):
    book_cases.kat1(emu)


@pytest.mark.parametrize("fixture,n_ops", [("env_IBM_2003-01-14_s789.npz", 10000), ("ddqn_IBM_2003-01-14_s4242.npz", 12000)])
def test_recorded_operation_tape(emu, golden_dir, fixture, n_ops):
    fills, modifies = book_cases.recorded_tape(golden_dir, fixture, n_ops, emu)
    assert fills > 500 and modifies > 1000


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_adversarial_tape(emu, seed):
    mods, execs = book_cases.random_tape_vs_oracle(emu, seed=seed)
    assert mods > 200 and execs > 500


@pytest.mark.parametrize("seed,reprice", [(0, 0.3), (1, 0.6), (2, 1.0), (3, 0.1), (4, 0.5), (5, 0.8)])
def test_adversarial_tape_with_repricing_modifies(emu, seed, reprice):
    mods, execs = book_cases.random_tape_vs_oracle(emu, seed=seed, reprice=reprice)
    assert mods > 100 and execs > 300


def test_repricing_modifies_match_the_live_reference(emu, golden_dir):
    assert book_cases.reprice_golden(golden_dir, emu) > 300


def test_reference_method_names(emu):
    """KAT-1 spelled with the reference's own calls: OrderBook.handleLimitOrder / modifyOrder / cancelOrder / getInsideBids /
    getInsideAsks / last_trade (util/OrderBook.py:38,284,341,377-398)."""
    from marl_optimal_execution_b200.book import LimitOrder, OrderBookBatch
    b = OrderBookBatch(n_envs=2, trace_cap=256, level_cap=64, order_cap=64, lib_path=emu)
    b.currentTime = book_cases.T0
    o12 = LimitOrder(2, b.currentTime, "JPM", 200, True, 1000, order_id=12)
    for o in (LimitOrder(1, b.currentTime, "JPM", 100, True, 1000, order_id=11), o12, LimitOrder(3, b.currentTime, "JPM", 300, True, 999, order_id=13),
              LimitOrder(4, b.currentTime, "JPM", 50, False, 1005, order_id=14), LimitOrder(5, b.currentTime, "JPM", 60, False, 1003, order_id=15)):
        b.handleLimitOrder(o)
    assert b.getInsideBids(2) == [(1000, 300), (999, 300)] and b.getInsideAsks(env=1) == [(1003, 60), (1005, 50)] and b.last_trade is None
    b.modifyOrder(o12, LimitOrder(2, b.currentTime, "JPM", 150, True, 1000, order_id=12))
    b.modifyOrder(o12, LimitOrder(2, b.currentTime, "JPM", 1, True, 1000, order_id=77))            # not the same order: ignored
    assert b.getInsideBids(1) == [(1000, 350)]                                                     # slot 0 overwritten: 150 + 200
    b.cancelOrder(LimitOrder(9, b.currentTime, "JPM", 1, True, 1000, order_id=99))                 # unknown id: silent no-op
    b.handleLimitOrder(LimitOrder(6, b.currentTime, "JPM", 400, False, 999, order_id=16))
    assert b.getInsideBids() == [(999, 250)] and b.getInsideAsks() == [(1003, 60), (1005, 50)] and b.last_trade == 1000
    notes, _ = b.notifications(0)
    assert [int(r[2]) for r in notes].count(book_cases.EXEC) == 6 and [int(r[2]) for r in notes].count(book_cases.MOD) == 1
    b.close()
