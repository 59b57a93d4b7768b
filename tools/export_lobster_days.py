#!/usr/bin/env python3
"""Export replayed order streams of LOBSTER sample days as small fixtures (build container only; the GPU box has no /root/reference):

    python tools/export_lobster_days.py IBM 2003-01-13 2003-01-17 2003-01-21 ...  ->  tests/golden/days/IBM_<date>.npz  (key "stream": int64 [n, 5])

Rows are (t_ns since midnight, ORDER_ID, PRICE cents, SIZE, is_buy) exactly as LOBSTEROrdersProcessor.processOrders builds them
(agent/examples/MarketReplayAgent.py:196-216); marl_optimal_execution_b200.env.load_lobster_csv is pinned to the reference's own parse of the IBM / GOOG
sample days by tests/test_lobster_loader.py.  The nine train dates of config/execution/marketreplay/execution_marketreplay_ddqn_parallel.py:40-44 are
2003-01-13..17 and 2003-01-21..24."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_optimal_execution_b200.env import load_lobster_csv, lobster_message_path   # noqa: E402

ticker, dates = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "tests", "golden", "days")
os.makedirs(out_dir, exist_ok=True)
for d in dates:
    rows = load_lobster_csv(lobster_message_path(ticker, d, "/root/reference/data/lobster"))
    assert len(rows) and rows[:, 2].max() < 2 ** 31 and rows[:, 3].max() < 2 ** 31
    np.savez_compressed(os.path.join(out_dir, "%s_%s.npz" % (ticker, d)), stream=rows)
    print(d, len(rows), "rows,", len(np.unique(rows[:, 0])), "timestamps")
