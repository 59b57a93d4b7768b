"""OrderBookBatch: the reference's book surface (util/OrderBook.py as ExchangeAgent drives it, agent/ExchangeAgent.py:311,324,339) for
n_envs books at once, fed by an operation tape -- e.g. the rows tools/record_reference.py records at the reference's exchange boundary.

    books = OrderBookBatch(n_envs=4096, trace_cap=100000)
    books.replay(ops)                     # int64 [n, 9]: (t_ns, op 0 limit / 1 cancel / 2 modify, agent, order_id, is_buy, price, qty, new_price, new_qty)
    notes, snaps = books.notifications(0) # what the book sent to its owner, and the book state after every operation
    books.inside(0, is_bid=True, depth=5) # getInsideBids(5)
"""
import ctypes as C

import numpy as np

from . import _lib


class LimitOrder:
    """The fields of util/order/LimitOrder.py:14-20 (+ Order.py:10-33) the book reads."""

    def __init__(self, agent_id, time_placed, symbol, quantity, is_buy_order, limit_price, order_id=None):
        self.agent_id, self.time_placed, self.symbol, self.quantity = agent_id, time_placed, symbol, quantity
        self.is_buy_order, self.limit_price, self.order_id, self.fill_price = is_buy_order, limit_price, order_id, None


class OrderBookBatch:
    def __init__(self, n_envs=1, stream_history=10, level_cap=1024, order_cap=16384, trace_cap=0, device=0, lib_path=None):
        self._L = _lib.load(lib_path)
        self.n_envs, self.trace_cap = int(n_envs), int(trace_cap)
        self.level_cap = int(level_cap)
        self.currentTime = 0          # owner.currentTime (ns) stamped on the operations of the per-call methods below
        self._h = C.c_void_p()
        _lib.check(self._L, self._L.abx_book_create(int(stream_history), int(level_cap), int(order_cap), int(trace_cap), self.n_envs, int(device), C.byref(self._h)),
                   "abx_book_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.abx_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def replay(self, ops, stream=None):
        o = np.ascontiguousarray(ops, dtype=np.int64)
        if o.ndim != 2 or o.shape[1] != 9:
            raise ValueError("ops must be int64 [n, 9]")
        _lib.check(self._L, self._L.abx_book_replay(self._h, o.ctypes.data_as(C.POINTER(C.c_int64)), len(o), stream), "abx_book_replay")

    def stats(self, stream=None):
        out = np.zeros(self.n_envs, dtype=_lib.STATS_DTYPE)
        _lib.check(self._L, self._L.abx_sim_stats(self._h, out.ctypes.data, stream), "abx_sim_stats")
        return out

    def notifications(self, env, stream=None):
        """(notes int64 [k, 13], snaps int64 [n_ops, 16]) in the row layouts of tools/record_reference.py."""
        out = np.zeros(max(self.trace_cap, 1), dtype=_lib.TRACE_DTYPE)
        n = C.c_int32(0)
        _lib.check(self._L, self._L.abx_sim_trace(self._h, int(env), out.ctypes.data, self.trace_cap, C.byref(n), stream), "abx_sim_trace")
        tr = out[: n.value]
        a, b = tr[tr["tag"] == 1], tr[tr["tag"] == 2]
        notes = np.zeros((len(a), 13), dtype=np.int64)
        notes[:, 0], notes[:, 1], notes[:, 2:] = a["t"], a["a"], a["v"][:, :11]
        return notes, b["v"].astype(np.int64)

    def inside(self, env, is_bid, depth, stream=None):
        out = np.zeros(2 * max(depth, 1), dtype=np.int32)
        n = C.c_int32(0)
        _lib.check(self._L, self._L.abx_sim_book_snapshot(self._h, int(env), int(bool(is_bid)), int(depth), out.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n), stream),
                   "abx_sim_book_snapshot")
        return [(int(out[2 * k]), int(out[2 * k + 1])) for k in range(n.value)]

    # ---- the reference's method names (util/OrderBook.py:38,284,341,377-398), one operation applied to every book of the batch; a thin
    #      host-side spelling of replay() for code written against the reference -- bulk work belongs in one replay() tape
    def _one(self, kind, o, new=None):
        if o.order_id is None:
            raise ValueError("order_id must be set (the exchange path generates ids, util/order/Order.py:27; the bare book does not)")
        self.replay(np.array([[int(self.currentTime), kind, int(o.agent_id), int(o.order_id), int(bool(o.is_buy_order)), int(o.limit_price), int(o.quantity),
                               int(new.limit_price) if new is not None else 0, int(new.quantity) if new is not None else 0]], dtype=np.int64))

    def handleLimitOrder(self, order):
        self._one(0, order)

    def cancelOrder(self, order):
        self._one(1, order)

    def modifyOrder(self, order, new_order):
        if new_order.order_id != order.order_id:          # isSameOrder, util/OrderBook.py:343,458-459: silently ignored
            return
        self._one(2, order, new_order)

    def getInsideBids(self, depth=None, env=0):
        return self.inside(env, True, self.level_cap if depth is None else min(int(depth), self.level_cap))

    def getInsideAsks(self, depth=None, env=0):
        return self.inside(env, False, self.level_cap if depth is None else min(int(depth), self.level_cap))

    @property
    def last_trade(self):
        """last_trade of book 0 (None before the first trade, util/OrderBook.py:24)."""
        lt = int(self.stats()["last_trade"][0])
        return None if lt < 0 else lt

