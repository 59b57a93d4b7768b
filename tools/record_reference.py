#!/usr/bin/env python3
"""Record golden traces from the UNMODIFIED reference (/root/reference) for tests/golden/.

Runs in the build container only (the reference is a Python program and cannot travel to the
GPU box).  Nothing from /root/reference is copied: this script imports it with the App. D shims
(tools/shims) and hooks four observation points at run time:

  * kernel event-queue pops      Kernel.py:192       -> (t_ns, recipient, type, uniq, msg kind)
  * book operations (inputs)     util/OrderBook.py:38,284,341 (called from agent/ExchangeAgent.py:311,324,339)
  * exchange outbound messages   agent/ExchangeAgent.py:471-485 -> fills, accepts, cancels, L1 replies
  * every RandomState draw       as the *standard* variate (SURVEY App. C identities) per stream

Usage:  python tools/record_reference.py sparse_zi_100 123456789 tests/golden/z100_s123456789.npz [--full]
        --full  also stores the pop / op / notification traces and RNG tapes (default stores
                checkpointed hashes, counts and final holdings only).
        python tools/record_reference.py marketreplay 1 tests/golden/mr_GOOG_2012-06-21.npz --date 2012-06-21 --extra -t GOOG -d 2012-06-21
        python tools/record_reference.py marketreplay 1 tests/golden/mr_sample_orders_file.npz --full --date 2019-06-03 \
               --orders-csv /root/reference/data/sample_orders_file.csv --extra -t SAMPLE -d 2019-06-03
        --orders-csv  replay a plain L3 order file (TIMESTAMP,ORDER_ID,PRICE,SIZE,BUY_SELL_FLAG) instead of a LOBSTER day
        --pov-exec POV QTY BUY|SELL  (rmsc03) append the reference's POVExecutionAgent to the config's agent list
        --exec-agent passive|aggressive HH:MM:SS QTY BUY|SELL [LIMIT]  (rmsc03) append the reference's PassiveAgent / AggressiveAgent
                (agent/execution/baselines/passive_agent.py, aggressive_agent.py: one limit / one market order at a timestamp; no shipped config instantiates them)
        --stop HH:MM:SS  replace Kernel.runner's stopTime (shortened recordings of long configs, e.g. rmsc01)
        python tools/record_reference.py rmsc01 123456789 tests/golden/rmsc01_s123456789_0945.npz --full --stop 09:45:00
"""
import importlib
import os
import queue
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ABIDES_REFERENCE", "/root/reference")

# Message kinds (the reference's msg.body["msg"] strings), in the numbering the
# C-ABI uses (include/abides_b200.h, enum abx_msg_kind).
MSG_KINDS = [
    "NONE", "WHEN_MKT_OPEN", "WHEN_MKT_CLOSE", "QUERY_SPREAD", "LIMIT_ORDER", "CANCEL_ORDER",
    "MODIFY_ORDER", "ORDER_ACCEPTED", "ORDER_EXECUTED", "ORDER_CANCELLED", "MKT_CLOSED",
    "QUERY_LAST_TRADE", "QUERY_TRANSACTED_VOLUME", "ORDER_MODIFIED", "QUERY_ORDER_STREAM", "MARKET_DATA",
    "MARKET_DATA_SUBSCRIPTION_REQUEST", "MARKET_DATA_SUBSCRIPTION_CANCELLATION",
]
KIND = {k: i for i, k in enumerate(MSG_KINDS)}

FNV_OFF = 0xCBF29CE484222325
FNV_PRIME = 0x100000001B3
M64 = (1 << 64) - 1


def fnv_mix(h, v):
    """FNV-1a over the 8 little-endian bytes of v (two's complement)."""
    v &= M64
    for _ in range(8):
        h = ((h ^ (v & 0xFF)) * FNV_PRIME) & M64
        v >>= 8
    return h


class Recorder:
    def __init__(self):
        self.midnight = None
        self.pops = []       # (t, recipient, type, uniq, kind)
        self.ops = []        # (t, op, agent, order_id, is_buy, price, qty, new_price, new_qty)
        self.notes = []      # (t, recipient, kind, order_id, is_buy, qty, price, fill_price|last_trade, bid, bid_q, ask, ask_q, mkt_closed)
        self.snaps = []      # after every book op: (n_bid_lv, n_ask_lv, n_resting, b0,bq0,b1,bq1,b2,bq2, a0,aq0,a1,aq1,a2,aq2, last_trade)
        self.streams = []    # RecRS objects in creation order
        self.global_tape = []
        self.runtime = False       # set when Kernel.runner starts: global-stream draws after that point are runtime draws
        self.gkind, self.gbits = [], []   # runtime draws on the GLOBAL np.random stream (kinds 'e','u','i'), in order

    def ns(self, ts):
        return int((ts - self.midnight).value)


REC = Recorder()


class RecRS(np.random.RandomState):
    """RandomState that logs each draw as its standard variate (bit-equal identities)."""

    def __init__(self, seed=None):
        super().__init__(seed)
        self.tape_kind = []
        self.tape_val = []
        self.seed_value = None if seed is None else int(seed)
        REC.streams.append(self)

    def _rec(self, k, v):
        self.tape_kind.append(k)
        self.tape_val.append(v)

    def normal(self, loc=0.0, scale=1.0, size=None):
        if size is None:
            z = super().standard_normal()
            self._rec(b"n", z)
            return loc + scale * z
        n = int(np.prod(size))
        out = np.empty(n)
        for i in range(n):
            z = super().standard_normal()
            self._rec(b"n", z)
            out[i] = loc + scale * z
        return out.reshape(size)

    def exponential(self, scale=1.0, size=None):
        assert size is None
        e = super().standard_exponential()
        self._rec(b"e", e)
        return e * scale

    def uniform(self, low=0.0, high=1.0, size=None):
        assert size is None
        u = super().random_sample()
        self._rec(b"u", u)
        return low + (high - low) * u

    def randint(self, low, high=None, size=None, dtype=int):
        v = super().randint(low, high, size=size, dtype=dtype)
        lo = 0 if high is None else low
        if size is None:
            self._rec(b"i", int(v) - int(lo))
        else:
            for x in np.asarray(v).ravel():
                self._rec(b"i", int(x) - int(lo))
        return v

    def choice(self, a, size=None, replace=True, p=None):
        # Kernel.py:411 passes the noise list as `replace`; numpy then draws randint(0, a, size).
        assert isinstance(a, int) and p is None
        v = super().randint(0, a, size=size)
        for x in np.asarray(v).ravel():
            self._rec(b"i", int(x))
        return v


class RecPQ(queue.PriorityQueue):
    def get(self, *a, **k):
        item = super().get(*a, **k)
        t, (recipient, mtype, msg) = item
        if REC.midnight is not None:
            if msg is None:
                uniq, kind = -1, 0
            else:
                uniq, kind = msg.uniq, KIND[msg.body["msg"]]
            REC.pops.append((REC.ns(t), int(recipient), int(mtype.value), uniq, kind))
        return item


def install_hooks():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "shims"))
    import pandas
    import pandas.io.json

    if not hasattr(pandas.io.json, "json_normalize"):  # same alias tools/shims/sitecustomize.py installs
        pandas.io.json.json_normalize = pandas.json_normalize

    if not hasattr(pandas.Timedelta, "delta"):       # pandas 2 dropped Timedelta.delta (total nanoseconds); agent/ExchangeAgent.py:373 and the subscription agents use it
        pandas.Timedelta.delta = property(lambda self: self.value)
    queue.PriorityQueue = RecPQ
    np.random.RandomState = RecRS

    # Global-stream draws that happen at run time (megashock gaps, SparseMeanRevertingOracle.py:69,168).
    g_exponential = np.random.exponential

    g_state = np.random.mtrand._rand

    def rec_exponential(scale=1.0, size=None):
        e = g_state.standard_exponential()
        REC.global_tape.append(float(e))
        if REC.runtime:
            REC.gkind.append(b"e"); REC.gbits.append(np.float64(e).view(np.uint64))
        return e * scale

    np.random.exponential = rec_exponential
    g_rand, g_randint = np.random.rand, np.random.randint

    def rec_rand(*shape):
        v = g_rand(*shape)
        if REC.runtime and not shape:
            REC.gkind.append(b"u"); REC.gbits.append(np.float64(v).view(np.uint64))
        return v

    def rec_randint(low, high=None, size=None, dtype=int):
        v = g_randint(low, high, size=size, dtype=dtype)
        if REC.runtime and size is None:
            REC.gkind.append(b"i"); REC.gbits.append(np.uint64(int(v) - (0 if high is None else int(low))))
        return v

    np.random.rand = rec_rand
    np.random.randint = rec_randint
    import Kernel as K
    r0 = K.Kernel.runner

    def runner(self, *a, **k):
        REC.runtime = True
        return r0(self, *a, **k)

    K.Kernel.runner = runner

    import util.OrderBook as OB
    import agent.ExchangeAgent as EA

    def snap(book):
        b = book.getInsideBids(3)
        a = book.getInsideAsks(3)
        row = [len(book.bids), len(book.asks), sum(len(l) for l in book.bids) + sum(len(l) for l in book.asks)]
        for side in (b, a):
            for i in range(3):
                row += list(side[i]) if i < len(side) else [0, 0]
        row.append(-1 if book.last_trade is None else int(book.last_trade))
        REC.snaps.append(tuple(int(x) for x in row))

    h0, c0, m0 = OB.OrderBook.handleLimitOrder, OB.OrderBook.cancelOrder, OB.OrderBook.modifyOrder

    def handleLimitOrder(self, order):
        REC.ops.append((REC.ns(self.owner.currentTime), 0, order.agent_id, order.order_id, int(order.is_buy_order),
                        int(order.limit_price), int(order.quantity), 0, 0))
        h0(self, order)
        snap(self)

    def cancelOrder(self, order):
        REC.ops.append((REC.ns(self.owner.currentTime), 1, order.agent_id, order.order_id, int(order.is_buy_order),
                        int(order.limit_price), int(order.quantity), 0, 0))
        c0(self, order)
        snap(self)

    def modifyOrder(self, order, new_order):
        REC.ops.append((REC.ns(self.owner.currentTime), 2, order.agent_id, order.order_id, int(order.is_buy_order),
                        int(order.limit_price), int(order.quantity), int(new_order.limit_price),
                        int(new_order.quantity)))
        m0(self, order, new_order)
        snap(self)

    OB.OrderBook.handleLimitOrder = handleLimitOrder
    OB.OrderBook.cancelOrder = cancelOrder
    OB.OrderBook.modifyOrder = modifyOrder

    s0 = EA.ExchangeAgent.sendMessage

    def sendMessage(self, recipientID, msg):
        b = msg.body
        kind = KIND[b["msg"]]
        row = [REC.ns(self.currentTime), int(recipientID), kind, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]
        o = b.get("order", b.get("new_order"))
        if o is not None:
            row[3:8] = [o.order_id, int(o.is_buy_order), int(o.quantity), int(o.limit_price),
                        0 if o.fill_price is None else int(o.fill_price)]
        if b["msg"] == "QUERY_SPREAD":
            row[7] = -1 if b["data"] is None else int(b["data"])      # last_trade None before the first trade (no oracle)
            if b["bids"]:
                row[8], row[9] = b["bids"][0]
            if b["asks"]:
                row[10], row[11] = b["asks"][0]
            row[12] = int(bool(b["mkt_closed"]))
        if b["msg"] == "MARKET_DATA":                                  # agent/ExchangeAgent.py:371-384: `levels` levels a side + last trade
            bids, asks = b["bids"], b["asks"]
            row[3], row[4] = len(bids), len(asks)
            row[5] = sum((i + 1) * int(p) for i, (p, q) in enumerate(bids))     # position-weighted sums: every level's price and size is pinned
            row[6] = sum((i + 1) * int(p) for i, (p, q) in enumerate(asks))
            row[7] = -1 if b["last_transaction"] is None else int(b["last_transaction"])
            if bids:
                row[8], row[9] = bids[0]
            if asks:
                row[10], row[11] = asks[0]
            row[12] = sum((i + 1) * (int(q) + 3 * int(q2)) for i, ((p, q), (p2, q2)) in enumerate(zip(bids, asks))) if bids and asks else 0
        REC.notes.append(tuple(int(x) for x in row))
        s0(self, recipientID, msg)

    EA.ExchangeAgent.sendMessage = sendMessage


def run(config, seed, extra_args=(), date="2019-06-28"):
    import pandas as pd

    REC.midnight = pd.to_datetime(date)
    sys.argv = ["abides.py", "-c", config, "-l", "rec", "-s", str(seed), *extra_args]
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="abides_rec_")
    os.makedirs(os.path.join(tmp, "data", "marketreplay", "level_1"))          # fresh cache: the CURRENT loader builds int-cent prices
    os.symlink(os.path.join(REF, "data", "lobster"), os.path.join(tmp, "data", "lobster"))
    os.chdir(tmp)
    try:
        mod = importlib.import_module("config." + config)
    finally:
        os.chdir(cwd)
    return mod


def main():
    config, seed, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rest = sys.argv[4:]
    full = "--full" in rest
    date = rest[rest.index("--date") + 1] if "--date" in rest else "2019-06-28"
    extra = rest[rest.index("--extra") + 1:] if "--extra" in rest else []
    install_hooks()
    if config == "rmsc03":
        # the shipped config calls a method that does not exist: one-line alias so that it runs
        import agent.TradingAgent as TA
        TA.TradingAgent.getTransactedVolume = TA.TradingAgent.get_transacted_volume
        import agent.ExchangeAgent as EA
        EA.ExchangeAgent.logOrderBookSnapshots = lambda self, symbol: None      # archival: pd.SparseDataFrame, out of scope
    if config in ("rmsc01", "rmsc02"):
        import agent.ExchangeAgent as EA
        EA.ExchangeAgent.logOrderBookSnapshots = lambda self, symbol: None      # archival after the run (book_freq "M" is rejected by current pandas): out of scope
    if config == "marketreplay":
        # order-book archival (agent/ExchangeAgent.py:389-469) uses pd.SparseDataFrame, removed from pandas; it runs after the
        # simulation ended and is out of scope, so the recorder skips it.
        import agent.ExchangeAgent as EA
        EA.ExchangeAgent.logOrderBookSnapshots = lambda self, symbol: None
    if "--orders-csv" in rest:
        # Replay a plain L3 order file (TIMESTAMP,ORDER_ID,PRICE,SIZE,BUY_SELL_FLAG; data/sample_orders_file.csv) instead of a LOBSTER day: the
        # reference's own L3OrdersProcessor cannot read that file (it expects a '|'-separated 17-column export,
        # agent/examples/MarketReplayAgent.py:140-147), so the recorder hands MarketReplayAgent the orders dict in the processor's output format
        # ({Timestamp: [{ORDER_ID, PRICE (cents = PRICE * 100, :152-153), SIZE, BUY_SELL_FLAG}]}, :158) and leaves everything after it unmodified.
        import pandas as pd
        import agent.examples.MarketReplayAgent as MRA
        csv_path = rest[rest.index("--orders-csv") + 1]

        def process_orders(self):
            df = pd.read_csv(csv_path, dtype=str)
            od = {}
            for r in df.itertuples(index=False):
                ts = pd.Timestamp(pd.to_datetime(r.TIMESTAMP[:21], format="%Y%m%d%H%M%S.%f"))      # microseconds, like convertDate (:126-130)
                od.setdefault(ts, []).append({"ORDER_ID": int(r.ORDER_ID), "PRICE": int(float(r.PRICE) * 100), "SIZE": int(r.SIZE), "BUY_SELL_FLAG": r.BUY_SELL_FLAG})
            return od

        MRA.LOBSTEROrdersProcessor.processOrders = process_orders
    pov_args = None
    if "--pov-exec" in rest:
        # "rmsc03 with a POV execution agent" (BASELINE.json configs[2]): the shipped config/rmsc03.py has no execution agent, so the recorder
        # appends the reference's own POVExecutionAgent (agent/execution/baselines/pov_agent.py, parameters in the style of
        # config/execution_iabs_plots.py:200-226 scaled to the 15-minute session) to the agent list the UNMODIFIED config script hands to
        # Kernel.runner.  The agent gets a fixed-seed RandomState (it never draws), so the config's own seed cascade is untouched.
        import Kernel as K
        from agent.execution.baselines.pov_agent import POVExecutionAgent
        i = rest.index("--pov-exec")
        pov, qty, direction = float(rest[i + 1]), int(rest[i + 2]), rest[i + 3]
        pov_args = (pov, qty, direction)
        r1 = K.Kernel.runner

        def runner_with_pov(self, agents, *a, **k):
            import warnings
            import pandas as pd
            day = pd.to_datetime(date)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ag = POVExecutionAgent(id=len(agents), name="POV_EXECUTION_AGENT", type="ExecutionAgent", symbol=agents[1].symbol, starting_cash=10000000,
                                       direction=direction, quantity=qty, pov=pov, start_time=day + pd.to_timedelta("09:32:00"), freq="30s",
                                       lookback_period="30s", end_time=day + pd.to_timedelta("09:43:00"), trade=True, log_orders=False,
                                       random_state=np.random.mtrand.RandomState.__new__(np.random.mtrand.RandomState))
            agents.append(ag)
            n = len(agents)
            k["agentLatency"] = np.zeros((n, n))
            return r1(self, agents=agents, *a, **k)

        K.Kernel.runner = runner_with_pov
    ex_args = None
    if "--exec-agent" in rest:
        import Kernel as K
        from agent.execution.baselines.passive_agent import PassiveAgent
        from agent.execution.baselines.aggressive_agent import AggressiveAgent
        i = rest.index("--exec-agent")
        kind, hms, qty, direction = rest[i + 1], rest[i + 2], int(rest[i + 3]), rest[i + 4]
        limit = int(rest[i + 5]) if len(rest) > i + 5 and rest[i + 5].lstrip("-").isdigit() else None
        ex_args = (kind, hms, qty, direction, limit)
        r2 = K.Kernel.runner

        def runner_with_exec(self, agents, *a, **k):
            import pandas as pd
            ts = pd.to_datetime(date) + pd.to_timedelta(hms)
            rs = np.random.mtrand.RandomState.__new__(np.random.mtrand.RandomState)
            if kind == "passive":
                ag = PassiveAgent(id=len(agents), name="PASSIVE_AGENT", type="PassiveAgent", symbol=agents[1].symbol, starting_cash=10000000, timestamp=ts,
                                  direction=direction, quantity=qty, limit_price=limit, log_orders=False, random_state=rs)
            else:
                ag = AggressiveAgent(id=len(agents), name="AGGRESSIVE_AGENT", type="AggressiveAgent", symbol=agents[1].symbol, starting_cash=10000000, timestamp=ts,
                                     direction=direction, quantity=qty, log_orders=False, random_state=rs)
            agents.append(ag)
            n = len(agents)
            k["agentLatency"] = np.zeros((n, n))
            return r2(self, agents=agents, *a, **k)

        K.Kernel.runner = runner_with_exec
    if "--stop" in rest:
        # shortened run: the UNMODIFIED config script is executed as is, only Kernel.runner's stopTime is replaced (rmsc01 under the current code makes
        # ~2 M messages per day, dominated by the market maker's 20 quotes per second; a recorded prefix of the day pins the same logic)
        import Kernel as K
        import pandas as pd
        hms = rest[rest.index("--stop") + 1]
        r0 = K.Kernel.runner

        def runner_with_stop(self, *a, **k):
            k["stopTime"] = pd.to_datetime(date) + pd.to_timedelta(hms)
            return r0(self, *a, **k)

        K.Kernel.runner = runner_with_stop
    mod = run(config, seed, extra, date)

    pops = np.array(REC.pops, dtype=np.int64).reshape(-1, 5)
    ops = np.array(REC.ops, dtype=np.int64).reshape(-1, 9)
    notes = np.array(REC.notes, dtype=np.int64).reshape(-1, 13)
    snaps = np.array(REC.snaps, dtype=np.int64).reshape(-1, 16)

    # Checkpointed FNV-1a hash of the pop sequence (every 1000 pops + final).
    h = FNV_OFF
    ck = []
    for i, row in enumerate(REC.pops):
        for v in row[:4]:
            h = fnv_mix(h, v)
        if (i + 1) % 1000 == 0:
            ck.append(h)
    ck.append(h)
    hn = FNV_OFF
    for row in REC.notes:
        for v in row:
            hn = fnv_mix(hn, v)
    hs = FNV_OFF
    for row in REC.snaps:
        for v in row:
            hs = fnv_mix(hs, v)

    agents = mod.agents
    sym = getattr(mod, "symbol", "JPM")
    hold = []
    for a in agents[1:]:
        shares = int(a.holdings.get(sym, 0))
        cash = int(a.holdings["CASH"])
        lt = int(a.last_trade[sym]) if sym in a.last_trade else 0
        surplus = [e["Event"] for e in a.log if e["EventType"] == "FINAL_VALUATION"]
        hold.append((a.id, shares, cash, cash + shares * lt, int(surplus[-1]) if surplus else 0))
    hold = np.array(hold, dtype=np.int64)

    data = dict(
        config=np.array(config), seed=np.array(seed), n_pops=np.array(len(pops)),
        pop_hash_ckpt=np.array(ck, dtype=np.uint64), note_hash=np.array(hn, dtype=np.uint64),
        snap_hash=np.array(hs, dtype=np.uint64), n_ops=np.array(len(ops)), n_notes=np.array(len(notes)),
        holdings=hold, kind_counts=np.bincount(pops[:, 4], minlength=len(MSG_KINDS)),
        type_counts=np.bincount(pops[:, 2], minlength=4),
        stream_seeds=np.array([-1 if s.seed_value is None else s.seed_value for s in REC.streams], dtype=np.int64),
        stream_draws=np.array([len(s.tape_val) for s in REC.streams], dtype=np.int64),
        global_exp_tape=np.array(REC.global_tape, dtype=np.float64),
        global_kind=np.frombuffer(b"".join(REC.gkind), dtype="S1"), global_bits=np.array(REC.gbits, dtype=np.uint64),
        max_levels=np.array([snaps[:, 0].max() if len(snaps) else 0, snaps[:, 1].max() if len(snaps) else 0]),
        max_resting=np.array(snaps[:, 2].max() if len(snaps) else 0),
    )
    replay = [a for a in agents if type(a).__name__ == "MarketReplayAgent"]
    if replay:                                   # the replayed stream exactly as LOBSTEROrdersProcessor parsed it
        od = replay[0].historical_orders.orders_dict
        data["stream"] = np.array([(REC.ns(ts), int(r["ORDER_ID"]), int(r["PRICE"]), int(r["SIZE"]), 1 if r["BUY_SELL_FLAG"] == "BUY" else 0)
                                   for ts in od for r in od[ts]], dtype=np.int64)
    data["pops_head"], data["notes_head"], data["snaps_head"] = pops[:20000], notes[:20000], snaps[:10000]
    if pov_args is not None:
        pa = agents[-1]
        data["pov_exec"] = np.array([pov_args[0], pov_args[1], 1 if pov_args[2] == "BUY" else 0, pa.rem_quantity, len(pa.executed_orders), len(pa.orders)], dtype=np.float64)
        data["pov_ops"] = ops[ops[:, 2] == pa.id]
    if ex_args is not None:
        xa = agents[-1]
        import pandas as pd
        data["exec_agent"] = np.array([1 if ex_args[0] == "passive" else 2, int(pd.to_timedelta(ex_args[1]).value), ex_args[2], 1 if ex_args[3] == "BUY" else 0,
                                       0 if ex_args[4] is None else ex_args[4], len(xa.orders)], dtype=np.int64)
        data["exec_ops"] = ops[ops[:, 2] == xa.id]
    if full:
        kinds = np.frombuffer(b"".join(b"".join(s.tape_kind) for s in REC.streams), dtype="S1")
        vals = []
        for s in REC.streams:
            for k, v in zip(s.tape_kind, s.tape_val):
                if k == b"i":
                    vals.append(np.int64(v).view(np.uint64))
                else:
                    vals.append(np.float64(v).view(np.uint64))
        data.update(
            pops=pops, ops=ops, notes=notes, snaps=snaps,
            tape_kind=kinds, tape_bits=np.array(vals, dtype=np.uint64),
            tape_offsets=np.concatenate([[0], np.cumsum([len(s.tape_val) for s in REC.streams])]).astype(np.int64),
        )
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    np.savez_compressed(out, **data)
    print("recorded", config, "seed", seed, "pops", len(pops), "ops", len(ops), "notes", len(notes),
          "streams", len(REC.streams), "->", out)


if __name__ == "__main__":
    main()
