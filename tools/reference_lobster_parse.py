#!/usr/bin/env python3
"""Parse one LOBSTER message file with the UNMODIFIED reference's LOBSTEROrdersProcessor (agent/examples/MarketReplayAgent.py:162-220; build
container only, imported with the SURVEY App. D shims) and save the orders dict as int64 rows (t_ns since midnight, ORDER_ID, PRICE, SIZE,
is_buy) -- the checker for marl_optimal_execution_b200.env.load_lobster_csv (tests/test_lobster_loader.py).

  python tools/reference_lobster_parse.py <message_csv> <yyyy-mm-dd> <out.npy>
"""
import contextlib
import io
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, os.environ.get("ABIDES_REFERENCE", "/root/reference"))
import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
import pandas.io.json  # noqa: E402

if not hasattr(pandas.io.json, "json_normalize"):
    pandas.io.json.json_normalize = pd.json_normalize
import util.util as uu  # noqa: E402

uu.silent_mode = True
from agent.examples.MarketReplayAgent import LOBSTEROrdersProcessor  # noqa: E402

path, date, out = sys.argv[1], sys.argv[2], sys.argv[3]
d = pd.to_datetime(date)
with contextlib.redirect_stdout(io.StringIO()):
    p = LOBSTEROrdersProcessor("X", d, d + pd.to_timedelta("09:30:00"), d + pd.to_timedelta("16:00:00"), path, tempfile.mkdtemp() + "/")
od = p.orders_dict
rows = [((ts - d).value, int(r["ORDER_ID"]), int(r["PRICE"]), int(r["SIZE"]), 1 if r["BUY_SELL_FLAG"] == "BUY" else 0) for ts in od for r in od[ts]]
np.save(out, np.array(rows, dtype=np.int64))
