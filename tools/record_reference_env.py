#!/usr/bin/env python3
"""Record a golden ABIDESEnv episode from the UNMODIFIED reference (build container only).

  python tools/record_reference_env.py IBM 2003-01-14 789 tests/golden/env_IBM_2003-01-14_s789.npz [n_steps]

Imports /root/reference with the SURVEY App. D shims (jsons, pandas json_normalize alias, gym / IPython stubs,
pandas frequency alias 'S' -> 's'), drives ABIDESEnv.reset()/step(action) (ABIDESEnv.py:30-57) with a fixed, seeded
action sequence and records: per step (action, obs, done), kernel pops, exchange-boundary ops, exchange outbound
messages, book snapshots (same hooks and row layouts as tools/record_reference.py), plus the replayed order stream
exactly as the reference's LOBSTEROrdersProcessor produced it (agent/examples/MarketReplayAgent.py:162-220).
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import record_reference as R  # noqa: E402


def install_stubs():
    gym = types.ModuleType("gym")

    class Env:
        pass

    class Box:
        def __init__(self, low, high):
            self.low, self.high, self.shape = low, high, np.shape(low)

    gym.Env = Env
    gym.spaces = types.SimpleNamespace(Box=Box)
    sys.modules["gym"] = gym
    ip = types.ModuleType("IPython")
    disp = types.ModuleType("IPython.display")
    disp.clear_output = lambda wait=False: None
    ip.display = disp
    sys.modules["IPython"] = ip
    sys.modules["IPython.display"] = disp
    import pandas as pd
    dr0 = pd.date_range

    def date_range(*a, **k):
        if isinstance(k.get("freq"), str):
            k["freq"] = k["freq"].replace("S", "s")
        return dr0(*a, **k)

    pd.date_range = date_range


def main():
    ticker, date, seed, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    max_steps = int(sys.argv[5]) if len(sys.argv) > 5 else 10 ** 9
    install_stubs()
    R.install_hooks()
    import pandas as pd
    R.REC.midnight = pd.to_datetime(date)

    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="abides_env_")
    os.makedirs(os.path.join(tmp, "data", "marketreplay", "level_1"))
    os.symlink(os.path.join(R.REF, "data", "lobster"), os.path.join(tmp, "data", "lobster"))
    os.chdir(tmp)
    sink = io.StringIO()
    try:
        import util.util as uu
        from util.order import LimitOrder as LO
        uu.silent_mode = True
        LO.silent_mode = True
        from ABIDESEnv import ABIDESEnv
        with contextlib.redirect_stdout(sink):
            env = ABIDESEnv(ticker=ticker, date=date, seed=seed)
        rl = env.agents.agent_list[2]
        rl.freq = "30s"                                   # pandas 3 rejects "30S" in Timestamp.floor
        replay = env.agents.agent_list[1]
        # the replayed stream exactly as the reference parsed it
        rows = []
        for ts in [replay.historical_orders.first_wakeup] + list(replay.wakeup_times):
            pass
        od = replay.historical_orders.orders_dict
        for ts in od:
            for r in od[ts]:
                rows.append((R.REC.ns(ts), int(r["ORDER_ID"]), int(r["PRICE"]), int(r["SIZE"]), 1 if r["BUY_SELL_FLAG"] == "BUY" else 0))
        arng = np.random.RandomState(seed)
        actions, obs_l, done_l, pops_at = [], [], [], []
        step = 0
        done = 0
        while not done and step < max_steps:
            a = np.array([arng.uniform(0.0, 0.04), arng.uniform(0, 1), arng.uniform(0, 1)])
            if step % 7 == 3:
                a[1:] = 0.0                                # exercises the all-zero level split (dummy_rl_execution_agent.py:146-147)
            with contextlib.redirect_stdout(sink):
                obs, reward, done, _ = env.step(a)
            sink.truncate(0); sink.seek(0)
            assert reward is None
            actions.append(a)
            o = np.full(9, np.nan)
            o[: len(obs)] = np.asarray(obs, dtype=float)
            obs_l.append(o)
            done_l.append(done)
            pops_at.append(len(R.REC.pops))
            step += 1
    finally:
        os.chdir(cwd)

    pops = np.array(R.REC.pops, dtype=np.int64).reshape(-1, 5)
    ops = np.array(R.REC.ops, dtype=np.int64).reshape(-1, 9)
    notes = np.array(R.REC.notes, dtype=np.int64).reshape(-1, 13)
    snaps = np.array(R.REC.snaps, dtype=np.int64).reshape(-1, 16)
    h, ck = R.FNV_OFF, []
    for i, row in enumerate(R.REC.pops):
        for v in row[:4]:
            h = R.fnv_mix(h, v)
        if (i + 1) % 1000 == 0:
            ck.append(h)
    ck.append(h)
    hn = R.FNV_OFF
    for row in R.REC.notes:
        for v in row:
            hn = R.fnv_mix(hn, v)
    hs = R.FNV_OFF
    for row in R.REC.snaps:
        for v in row:
            hs = R.fnv_mix(hs, v)
    keep = 40000
    np.savez_compressed(
        out, ticker=np.array(ticker), date=np.array(date), seed=np.array(seed),
        stream=np.array(rows, dtype=np.int64), actions=np.array(actions), obs=np.array(obs_l), done=np.array(done_l),
        pops_at_step=np.array(pops_at, dtype=np.int64), n_pops=np.array(len(pops)), n_ops=np.array(len(ops)),
        n_notes=np.array(len(notes)), pop_hash_ckpt=np.array(ck, dtype=np.uint64), note_hash=np.array(hn, dtype=np.uint64),
        snap_hash=np.array(hs, dtype=np.uint64), pops_head=pops[:keep], ops_head=ops[: keep // 4], notes_head=notes[:keep // 2],
        snaps_head=snaps[: keep // 4], kind_counts=np.bincount(pops[:, 4], minlength=len(R.MSG_KINDS)),
        max_levels=np.array([snaps[:, 0].max(), snaps[:, 1].max()]), max_resting=np.array(snaps[:, 2].max()),
        rl_final=np.array([rl.rem_quantity, rl.holdings.get(ticker, 0), rl.holdings["CASH"], len(rl.executed_orders)], dtype=np.float64),
        replay_final=np.array([replay.holdings.get(ticker, 0), replay.holdings["CASH"], len(replay.orders)], dtype=np.float64),
    )
    print("recorded env", ticker, date, "steps", step, "pops", len(pops), "ops", len(ops), "notes", len(notes), "->", out)
    print("kind counts", dict(zip(R.MSG_KINDS, np.bincount(pops[:, 4], minlength=len(R.MSG_KINDS)))))
    print("rl final", rl.rem_quantity, rl.holdings, "executed", len(rl.executed_orders))


if __name__ == "__main__":
    main()
