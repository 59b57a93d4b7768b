#!/usr/bin/env python3
"""Golden for the book surface under RE-PRICING modifies, produced by the LIVE reference util/OrderBook.py (unmodified, imported from
/root/reference): the adversarial operation tapes of tests/book_cases.py (few order ids re-used across prices and sides; a share of the
MODIFY rows changes the price, which leaves the reference's level lists unsorted with several levels showing one price,
util/OrderBook.py:350-352,381,393) are driven through OrderBook.handleLimitOrder / cancelOrder / modifyOrder with a stub owner; every
notification the book sends and the book state after every operation are written to tests/golden/book_reprice_tapes.npz.

    PYTHONPATH=tools/shims:/root/reference python tools/record_reference_book_tape.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import pandas as pd                                                    # noqa: E402

KIND = {"ORDER_ACCEPTED": 7, "ORDER_EXECUTED": 8, "ORDER_CANCELLED": 9, "ORDER_MODIFIED": 13}


def make_tape(seed, n_ops, n_ids, reprice):
    """Same generator as tests/book_cases.random_tape_vs_oracle (kept in step by tests/test_oracle_book.py, which regenerates the rows)."""
    T0 = 34200 * 10 ** 9
    rs = np.random.RandomState(seed)
    ops, t = [], T0
    for _ in range(n_ops):
        t += int(rs.randint(0, 3))
        kind = rs.choice(3, p=[0.5, 0.15, 0.35])
        oid, is_buy = int(rs.randint(1, n_ids + 1)), int(rs.randint(0, 2))
        price = 1000 + int(rs.randint(-6, 7)) + (0 if is_buy else 2)
        qty = int(rs.randint(1, 60))
        new_price = price + int(rs.randint(-3, 4)) if rs.uniform() < reprice else price
        ops.append((t, int(kind), int(rs.randint(1, 9)), oid, is_buy, price, qty, new_price, int(rs.randint(1, 60))))
    return np.array(ops, dtype=np.int64)


def run_reference(ops):
    import util.util as U
    U.silent_mode = True
    from util.OrderBook import OrderBook
    from util.order.LimitOrder import LimitOrder
    LimitOrder.silent_mode = True
    midnight = pd.Timestamp("2019-06-28")

    class Owner:                                                        # what the book reads of its ExchangeAgent
        stream_history, book_freq, name = 10, None, "stub"

        def __init__(self):
            self.notes, self.currentTime = [], midnight

        def sendMessage(self, agent_id, msg):
            o = msg.body.get("order") or msg.body.get("new_order")
            self.notes.append((int((self.currentTime - midnight).value), agent_id, KIND[msg.body["msg"]], o.order_id, int(o.is_buy_order), o.quantity, o.limit_price,
                               o.fill_price if o.fill_price is not None else 0, 0, 0, 0, 0, 0))

        def logEvent(self, *a, **k):
            pass

    own = Owner()
    book = OrderBook(own, "X")
    snaps = []
    for t, kind, agent, oid, is_buy, price, qty, new_price, new_qty in ops.tolist():
        own.currentTime = midnight + pd.Timedelta(t, unit="ns")
        o = LimitOrder(agent, own.currentTime, "X", qty, bool(is_buy), price, order_id=oid)
        if kind == 0:
            book.handleLimitOrder(o)
        elif kind == 1:
            book.cancelOrder(o)
        else:
            book.modifyOrder(o, LimitOrder(agent, own.currentTime, "X", new_qty, bool(is_buy), new_price, order_id=oid))
        b3, a3 = book.getInsideBids(3), book.getInsideAsks(3)
        rest = sum(len(lv) for lv in book.bids) + sum(len(lv) for lv in book.asks)
        snaps.append((len(book.bids), len(book.asks), rest) + tuple(x for pq in (b3 + [(0, 0)] * 3)[:3] for x in pq) + tuple(x for pq in (a3 + [(0, 0)] * 3)[:3] for x in pq)
                     + (book.last_trade if book.last_trade is not None else -1,))
    return np.array(own.notes, dtype=np.int64).reshape(-1, 13), np.array(snaps, dtype=np.int64)


if __name__ == "__main__":
    out = {}
    for seed, reprice in ((0, 0.3), (2, 1.0), (4, 0.5)):
        ops = make_tape(seed, 2500, 40, reprice)
        notes, snaps = run_reference(ops)
        out["notes_s%d" % seed], out["snaps_s%d" % seed], out["params_s%d" % seed] = notes, snaps, np.array([seed, 2500, 40, int(reprice * 100)])
        print("seed %d reprice %.1f: %d notifications, max levels %d / %d, modified %d" % (seed, reprice, len(notes), snaps[:, 0].max(), snaps[:, 1].max(), int((notes[:, 2] == 13).sum())))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "book_reprice_tapes.npz"), **out)
