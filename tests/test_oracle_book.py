"""Known-answer test KAT-1 (produced by the live reference util/OrderBook.py) on the oracle book."""
from oracle.oracle import OracleBook

EXEC, ACC, CANC, MOD = 8, 7, 9, 13


def test_kat1():
    b = OracleBook(stream_history=10)
    b.set_time(34200 * 10 ** 9)
    b.limit(1, 11, True, 1000, 100)
    b.limit(2, 12, True, 1000, 200)
    b.limit(3, 13, True, 999, 300)
    b.limit(4, 14, False, 1005, 50)
    b.limit(5, 15, False, 1003, 60)
    assert b.inside(True, 2) == [(1000, 300), (999, 300)]
    assert b.inside(False, 2) == [(1003, 60), (1005, 50)]
    assert [tuple(r[2:8]) for r in b.take_notes()] == [
        (ACC, 11, 1, 100, 1000, 0), (ACC, 12, 1, 200, 1000, 0), (ACC, 13, 1, 300, 999, 0),
        (ACC, 14, 0, 50, 1005, 0), (ACC, 15, 0, 60, 1003, 0)]
    # MODIFY of the SECOND order in the level overwrites slot 0 (util/OrderBook.py:352)
    b.modify(2, 12, True, 1000, 1000, 150)
    assert b.level_orders(True, 0) == [(12, 150, 1000), (12, 200, 1000)]
    n = b.take_notes()
    assert len(n) == 1 and tuple(n[0][1:7]) == (2, MOD, 12, 1, 150, 1000)
    # unknown id: silent no-op
    b.cancel(9, 99, True, 1000)
    assert len(b.take_notes()) == 0 and b.n_resting() == 5
    # marketable sell walks the head of the best level only, fills at resting prices
    b.limit(6, 16, False, 999, 400)
    n = b.take_notes()
    assert [(int(r[1]), int(r[2]), int(r[3]), int(r[5]), int(r[7])) for r in n] == [
        (6, EXEC, 16, 150, 1000), (2, EXEC, 12, 150, 1000), (6, EXEC, 16, 200, 1000), (2, EXEC, 12, 200, 1000),
        (6, EXEC, 16, 50, 999), (3, EXEC, 13, 50, 999)]
    assert b.inside(True, 5) == [(999, 250)] and b.inside(False, 5) == [(1003, 60), (1005, 50)]
    assert b.last_trade == 1000                      # int(round(399950 / 400))


def test_cancel_partial_and_level_removal():
    b = OracleBook()
    b.limit(1, 1, False, 500, 100)
    b.limit(2, 2, False, 500, 70)
    b.limit(3, 3, True, 500, 30)                     # partial fill of the head
    assert b.inside(False, 1) == [(500, 140)]
    b.cancel(1, 1, False, 500)
    n = b.take_notes()
    assert tuple(n[-1][2:7]) == (CANC, 1, 0, 70, 500)   # ORDER_CANCELLED carries the book's remaining qty
    b.cancel(2, 2, False, 500)
    assert b.n_levels(False) == 0 and b.n_resting() == 0


def test_transacted_volume_dedup():
    b = OracleBook(stream_history=10)
    b.set_time(1000)
    b.limit(1, 1, False, 100, 50)
    b.limit(2, 2, True, 100, 50)                     # one fill: incoming records its pre-fill qty 50, resting records 50
    # both sides log (1000, 50): duplicates collapse (util/OrderBook.py:428)
    assert b.transacted_volume(10 ** 9) == 50


def test_repricing_modifies_match_the_live_reference(golden_dir):
    """Re-pricing MODIFY_ORDER (util/OrderBook.py:350-352: slot 0 takes the new order where the level stands; level prices are read from slot 0,
    :381,393; the lists go unsorted): the oracle book against three tapes recorded from the live reference OrderBook."""
    import book_cases
    assert book_cases.reprice_golden(golden_dir, use_oracle=True) > 300
