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


class OrderBookBatch:
    def __init__(self, n_envs=1, stream_history=10, level_cap=1024, order_cap=16384, trace_cap=0, device=0, lib_path=None):
        self._L = _lib.load(lib_path)
        self.n_envs, self.trace_cap = int(n_envs), int(trace_cap)
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
