"""Shared cases for the book surface (abx_book_* / OrderBookBatch): run by the CPU suite on the host emulation harness of the product
logic and by the GPU suite on the CUDA library."""
import os

import numpy as np

from marl_optimal_execution_b200.book import OrderBookBatch
from oracle.oracle import OracleBook

EXEC, ACC, CANC, MOD = 8, 7, 9, 13
T0 = 34200 * 10 ** 9


def op(t, kind, agent, oid, is_buy, price, qty, new_price=0, new_qty=0):
    return (t, kind, agent, oid, int(is_buy), price, qty, new_price, new_qty)


def kat1(lib_path=None):
    """known-answer vector produced by the live reference util/OrderBook.py (incl. the slot-0 overwrite of modifyOrder)."""
    ops = [op(T0, 0, 1, 11, 1, 1000, 100), op(T0, 0, 2, 12, 1, 1000, 200), op(T0, 0, 3, 13, 1, 999, 300), op(T0, 0, 4, 14, 0, 1005, 50), op(T0, 0, 5, 15, 0, 1003, 60)]
    b = OrderBookBatch(n_envs=3, trace_cap=256, level_cap=64, order_cap=64, lib_path=lib_path)
    b.replay(np.array(ops, dtype=np.int64))
    assert b.inside(0, True, 2) == [(1000, 300), (999, 300)] and b.inside(2, False, 2) == [(1003, 60), (1005, 50)]
    notes, snaps = b.notifications(1)
    assert [tuple(r[2:8]) for r in notes] == [(ACC, 11, 1, 100, 1000, 0), (ACC, 12, 1, 200, 1000, 0), (ACC, 13, 1, 300, 999, 0), (ACC, 14, 0, 50, 1005, 0), (ACC, 15, 0, 60, 1003, 0)]
    assert len(snaps) == 5 and tuple(snaps[-1][:3]) == (2, 2, 5) and snaps[-1][15] == -1                     # last_trade None so far
    b.replay(np.array([op(T0, 2, 2, 12, 1, 1000, 200, 1000, 150), op(T0, 1, 9, 99, 1, 1000, 1), op(T0, 0, 6, 16, 0, 999, 400)], dtype=np.int64))
    notes, snaps = b.notifications(1)
    new = notes[5:]
    assert tuple(new[0][1:7]) == (2, MOD, 12, 1, 150, 1000)                                                    # one ORDER_MODIFIED; the unknown-id cancel is silent
    assert [(int(r[1]), int(r[2]), int(r[3]), int(r[5]), int(r[7])) for r in new[1:]] == [
        (6, EXEC, 16, 150, 1000), (2, EXEC, 12, 150, 1000), (6, EXEC, 16, 200, 1000), (2, EXEC, 12, 200, 1000), (6, EXEC, 16, 50, 999), (3, EXEC, 13, 50, 999)]
    assert b.inside(0, True, 5) == [(999, 250)] and b.inside(0, False, 5) == [(1003, 60), (1005, 50)]
    st = b.stats()
    assert (st["last_trade"] == 1000).all() and (st["fills"] == 3).all() and (st["n_resting"] == 3).all()
    assert snaps[5][3:7].tolist() == [1000, 350, 999, 300]                                                     # after the modify: slot 0 overwritten (150 + 200)
    b.close()


def recorded_tape(golden_dir, fixture, n_ops, lib_path=None, n_envs=2):
    """An operation tape recorded at the reference's exchange boundary, replayed through the bare books: the book state after every
    operation must equal the reference's own (snaps recorded from util/OrderBook.py), and every notification the oracle book's."""
    g = np.load(os.path.join(golden_dir, fixture))
    ops = g["ops_head"][:n_ops]
    b = OrderBookBatch(n_envs=n_envs, trace_cap=8 * n_ops, level_cap=1024, order_cap=32768, lib_path=lib_path)
    half = n_ops // 2
    b.replay(ops[:half]); b.replay(ops[half:])                                                               # a second call continues on the live books
    notes, snaps = b.notifications(n_envs - 1)
    assert np.array_equal(snaps, g["snaps_head"][:n_ops])                                                      # vs the live reference
    o = OracleBook(stream_history=10)
    for r in ops:
        o.set_time(int(r[0]))
        if r[1] == 0:
            o.limit(int(r[2]), int(r[3]), bool(r[4]), int(r[5]), int(r[6]))
        elif r[1] == 1:
            o.cancel(int(r[2]), int(r[3]), bool(r[4]), int(r[5]))
        else:
            o.modify(int(r[2]), int(r[3]), bool(r[4]), int(r[5]), int(r[7]), int(r[8]))
    ref = o.take_notes()
    assert notes.shape == ref.shape and np.array_equal(notes[:, :8], ref[:, :8]), (notes.shape, ref.shape)
    assert (b.stats()["flags"] == 1).all()
    assert b.inside(0, True, 3) == o.inside(True, 3) and b.inside(0, False, 3) == o.inside(False, 3)
    b.close()
    return int((ref[:, 2] == EXEC).sum()) // 2, int((ops[:, 1] == 2).sum())


def random_tape_vs_oracle(lib_path=None, seed=0, n_ops=6000, n_ids=40, reprice=0.0):
    """Adversarial tape: few order ids re-used across prices and sides, so that head-slot copies (modifyOrder's slot-0 overwrite), ids resting
    in several levels and modifies / cancels at stale prices all occur; with reprice > 0 a share of the modifies changes the price, which
    leaves the reference's level lists unsorted and several levels showing one price (util/OrderBook.py:350-352,381,393); every notification and the book after every operation must equal
    the oracle book's (which is pinned to the reference)."""
    rs = np.random.RandomState(seed)
    ops, t = [], T0
    for _ in range(n_ops):
        t += int(rs.randint(0, 3))
        kind = rs.choice(3, p=[0.5, 0.15, 0.35])
        oid, is_buy = int(rs.randint(1, n_ids + 1)), int(rs.randint(0, 2))
        price = 1000 + int(rs.randint(-6, 7)) + (0 if is_buy else 2)
        qty = int(rs.randint(1, 60))
        new_price = price + int(rs.randint(-3, 4)) if rs.uniform() < reprice else price       # a re-pricing MODIFY: the head slot takes the new price where it stands
        ops.append(op(t, int(kind), int(rs.randint(1, 9)), oid, is_buy, price, qty, new_price, int(rs.randint(1, 60))))
    ops = np.array(ops, dtype=np.int64)
    b = OrderBookBatch(n_envs=2, trace_cap=16 * n_ops, level_cap=64 if reprice == 0 else 512, order_cap=8192, lib_path=lib_path)       # unsorted level lists keep many more levels alive
    b.replay(ops)
    notes, snaps = b.notifications(1)
    o = OracleBook(stream_history=10)
    ref_snaps = []
    for r in ops:
        o.set_time(int(r[0]))
        if r[1] == 0:
            o.limit(int(r[2]), int(r[3]), bool(r[4]), int(r[5]), int(r[6]))
        elif r[1] == 1:
            o.cancel(int(r[2]), int(r[3]), bool(r[4]), int(r[5]))
        else:
            o.modify(int(r[2]), int(r[3]), bool(r[4]), int(r[5]), int(r[7]), int(r[8]))
        ref_snaps.append((o.n_levels(True), o.n_levels(False), o.n_resting()) + tuple(x for pq in (o.inside(True, 3) + [(0, 0)] * 3)[:3] for x in pq)
                         + tuple(x for pq in (o.inside(False, 3) + [(0, 0)] * 3)[:3] for x in pq) + (o.last_trade if o.last_trade is not None else -1,))
    ref = o.take_notes()
    assert notes.shape == ref.shape and np.array_equal(notes[:, :8], ref[:, :8]), (notes.shape, ref.shape)
    assert np.array_equal(snaps, np.array(ref_snaps, dtype=np.int64))
    assert (b.stats()["flags"] == 1).all()
    b.close()
    return int((ref[:, 2] == MOD).sum()), int((ref[:, 2] == EXEC).sum())


def make_tape(seed, n_ops, n_ids, reprice):
    """The generator of random_tape_vs_oracle, as tools/record_reference_book_tape.py uses it."""
    rs = np.random.RandomState(seed)
    ops, t = [], T0
    for _ in range(n_ops):
        t += int(rs.randint(0, 3))
        kind = rs.choice(3, p=[0.5, 0.15, 0.35])
        oid, is_buy = int(rs.randint(1, n_ids + 1)), int(rs.randint(0, 2))
        price = 1000 + int(rs.randint(-6, 7)) + (0 if is_buy else 2)
        qty = int(rs.randint(1, 60))
        new_price = price + int(rs.randint(-3, 4)) if rs.uniform() < reprice else price
        ops.append(op(t, int(kind), int(rs.randint(1, 9)), oid, is_buy, price, qty, new_price, int(rs.randint(1, 60))))
    return np.array(ops, dtype=np.int64)


def reprice_golden(golden_dir, lib_path=None, use_oracle=False):
    """tests/golden/book_reprice_tapes.npz: three adversarial tapes with re-pricing modifies run through the LIVE reference util/OrderBook.py
    (tools/record_reference_book_tape.py).  Every notification and the book after every operation must equal the reference's -- for the oracle
    book (use_oracle) or for the product's books (CUDA library / its CPU emulation)."""
    g = np.load(os.path.join(golden_dir, "book_reprice_tapes.npz"))
    total_mod = 0
    for seed in (0, 2, 4):
        _, n_ops, n_ids, rp = (int(x) for x in g["params_s%d" % seed])
        ops = make_tape(seed, n_ops, n_ids, rp / 100.0)
        ref_notes, ref_snaps = g["notes_s%d" % seed], g["snaps_s%d" % seed]
        if use_oracle:
            o = OracleBook(stream_history=10)
            snaps = []
            for r in ops:
                o.set_time(int(r[0]))
                if r[1] == 0:
                    o.limit(int(r[2]), int(r[3]), bool(r[4]), int(r[5]), int(r[6]))
                elif r[1] == 1:
                    o.cancel(int(r[2]), int(r[3]), bool(r[4]), int(r[5]))
                else:
                    o.modify(int(r[2]), int(r[3]), bool(r[4]), int(r[5]), int(r[7]), int(r[8]))
                snaps.append((o.n_levels(True), o.n_levels(False), o.n_resting()) + tuple(x for pq in (o.inside(True, 3) + [(0, 0)] * 3)[:3] for x in pq)
                             + tuple(x for pq in (o.inside(False, 3) + [(0, 0)] * 3)[:3] for x in pq) + (o.last_trade if o.last_trade is not None else -1,))
            notes, snaps = o.take_notes(), np.array(snaps, dtype=np.int64)
        else:
            b = OrderBookBatch(n_envs=2, trace_cap=16 * n_ops, level_cap=256, order_cap=8192, lib_path=lib_path)
            b.replay(ops)
            notes, snaps = b.notifications(1)
            assert (b.stats()["flags"] == 1).all()
            b.close()
        assert notes.shape == ref_notes.shape and np.array_equal(notes[:, :8], ref_notes[:, :8]), (seed, notes.shape, ref_notes.shape)
        assert np.array_equal(snaps, ref_snaps), seed
        total_mod += int((ref_notes[:, 2] == MOD).sum())
    return total_mod
