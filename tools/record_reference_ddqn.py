#!/usr/bin/env python3
"""Record a golden run of the DDQN execution config from the UNMODIFIED reference (build container only).

  python tools/record_reference_ddqn.py IBM 2003-01-14 4242 tests/golden/ddqn_IBM_2003-01-14_s4242.npz [policy_seed [BUY|SELL]]
  python tools/record_reference_ddqn.py IBM 2003-01-15 77 tests/golden/ddqn_vwap_IBM_2003-01-15_s77.npz 94 BUY --vwap
      --vwap: the config's baseline execution agent is the reference's VWAPExecutionAgent (agent/execution/baselines/vwap_agent.py) instead of its TWAP agent, with a
      U-shaped volume profile pickled for volume_profile_path (the only way the class can be constructed as shipped); the profile and the schedule it yields are stored

Runs config/execution/marketreplay/execution_marketreplay_ddqn.py (Exchange + MarketReplayAgent + 7 MomentumAgents +
TWAPExecutionAgent + DDQLearningExecutionAgent, BUY 500 000 shares from 10:00 over 330 min at "30s") in `test` mode with
the SURVEY App. D shims plus stand-ins for the two packages the config imports that this image lacks:

  * tensorflow.keras (Model / Dense / Dropout / RMSprop / SGD): `Model.predict` is replaced by a *policy tape* -- a seeded
    RandomState picks the argmax column -- so that the agent's unmodified `choose_action` (np.argmax of predict,
    ddqlearning_execution_agent.py:362-364) walks through all 24 actions; the network arithmetic itself is NOT pinned by
    this recording (TensorFlow is absent from the image).
  * matplotlib (imported for plots only).

Recorded: kernel pops, exchange-boundary ops, exchange outbound messages, book snapshots (hooks of tools/record_reference.py),
the replayed stream as the reference parsed it, the momentum agents' sizes, and from the DDQN agent after the run: the
chosen action per tick, `observation` (6 features), `experience` (s, a, s', r), `price_path`, `action_hist`,
`step_reward_hist`; from both execution agents the final holdings / remaining quantity.
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

POLICY = {"rs": None, "actions": [], "mlp": None, "q": []}


def numpy_mlp_forward(flat, dims, x):
    """fp32 forward of util/model/QNets.py NNModel_1 (Dense + ReLU hidden layers, linear output; Dropout is the identity in predict) from the
    flat parameter layout of marl_optimal_execution_b200/qnet.py (per layer W[out][in] then b[out])."""
    h, p = np.asarray(x, dtype=np.float32), 0
    n_layers = len(dims) - 1
    for l in range(n_layers):
        w = flat[p:p + dims[l] * dims[l + 1]].reshape(dims[l + 1], dims[l]); p += dims[l] * dims[l + 1]
        b = flat[p:p + dims[l + 1]]; p += dims[l + 1]
        h = h @ w.T + b
        if l + 1 < n_layers:
            h = np.maximum(h, np.float32(0))
    return h.astype(np.float32)


def install_stubs():
    tf = types.ModuleType("tensorflow")
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    optim = types.ModuleType("tensorflow.keras.optimizers")

    class Dense:
        def __init__(self, units, activation=None):
            self.units, self.activation = units, activation

    class Dropout:
        def __init__(self, rate):
            self.rate = rate

    class Model:
        def __init__(self, name=None):
            self.name = name

        def compile(self, **kw):
            pass

        def load_weights(self, path):
            pass

        def save_weights(self, path):
            pass

        def predict(self, x):                      # policy tape: one-hot on a seeded random column (or, with --mlp, a real fp32 network)
            x = np.asarray(x)
            if POLICY["mlp"] is not None:
                flat, dims = POLICY["mlp"]
                q = numpy_mlp_forward(flat, dims, x)
                for i in range(x.shape[0]):
                    POLICY["actions"].append(int(np.argmax(q[i]))); POLICY["q"].append(q[i].copy())
                return q
            q = np.zeros((x.shape[0], 24), dtype=np.float32)
            for i in range(x.shape[0]):
                a = int(POLICY["rs"].randint(0, 24))
                q[i, a] = 1.0
                POLICY["actions"].append(a)
            return q

    class _Opt:
        def __init__(self, *a, **k):
            pass

    keras.Model = Model
    layers.Dense, layers.Dropout = Dense, Dropout
    optim.RMSprop, optim.SGD = _Opt, _Opt
    keras.layers, keras.optimizers = layers, optim
    tf.keras = keras
    sys.modules.update({"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
                        "tensorflow.keras.optimizers": optim})
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})


def main():
    ticker, date, seed, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    pseed = int(sys.argv[5]) if len(sys.argv) > 5 else seed + 17
    direction = sys.argv[6] if len(sys.argv) > 6 else "BUY"
    if len(sys.argv) > 7 and sys.argv[7] == "--mlp":      # actions = argmax of a seeded Glorot-initialised network instead of the random tape
        sys.path.insert(0, os.path.dirname(HERE))
        from marl_optimal_execution_b200.qnet import DEFAULT_DIMS, init_params
        flat = init_params(DEFAULT_DIMS, seed=pseed)
        flat = (flat + np.random.RandomState.__new__(np.random.RandomState).__class__(pseed + 1).normal(0, 0.15, size=flat.shape)).astype(np.float32)   # trained-like: non-zero biases, no ties
        POLICY["mlp"] = (flat, DEFAULT_DIMS)
    vwap = "--vwap" in sys.argv
    install_stubs()
    POLICY["rs"] = np.random.RandomState.__new__(np.random.RandomState)
    np.random.RandomState.__init__(POLICY["rs"], pseed)          # created before the hooks: not one of the reference's streams
    R.install_hooks()
    import pandas as pd
    import agent.ExchangeAgent as EA
    EA.ExchangeAgent.logOrderBookSnapshots = lambda self, symbol: None      # archival (pd.SparseDataFrame), out of scope
    import agent.Agent as AG
    AG.Agent.writeLog = lambda self, dfLog, filename=None: None             # bz2 log files, out of scope
    R.REC.midnight = pd.to_datetime(date)
    if vwap:
        # The config script names TWAPExecutionAgent; the recorder swaps the NAME it imports for a factory of the reference's own VWAPExecutionAgent, so that the
        # unmodified config builds, registers and runs that class.  Profile: x^2 + 2x + 2 over the horizon's 30 s bins, normalised (the shape of the class's own
        # synthetic_volume_profile, which cannot be used: its freq must be an int for f"{freq}s" while interval_range needs a string / Timedelta).
        import pickle
        import agent.execution.baselines.twap_agent as TW
        from agent.execution.baselines.vwap_agent import VWAPExecutionAgent
        prof_path = os.path.join(tempfile.mkdtemp(prefix="abides_vwap_"), "profile.pkl")

        def make_vwap(**kw):
            kw.pop("log_events", None)                       # VWAPExecutionAgent.__init__ has no such parameter
            hz = kw["execution_time_horizon"]
            lefts = pd.date_range(hz[0], hz[-1], freq=kw["freq"])[:-1]
            n = len(lefts)
            w = np.array([x * x + 2 * x + 2 for x in range(int(-n / 2), int(-n / 2) + n)], dtype=np.float64)
            w = w * (1.0 / w.sum())
            ser = pd.Series(w, index=lefts)
            with open(prof_path, "wb") as f:
                pickle.dump(ser, f)
            POLICY["vwap_profile"] = w
            ag = VWAPExecutionAgent(volume_profile_path=prof_path, **kw)
            POLICY["vwap_schedule"] = np.array([ag.schedule[b] for b in ag.schedule], dtype=np.int64)
            return ag

        TW.TWAPExecutionAgent = make_vwap
    cfg = "execution.marketreplay.execution_marketreplay_ddqn"
    sys.argv = ["abides.py", "-c", cfg, "-s", str(seed), "-t", ticker, "-d", date, "--direction", direction, "--parent_qty", "500000",
                "--start_hour", "10", "--horizon_length", "330", "--freq", "30s", "-m", "test", "-a", "rl", "-l", "rec_ddqn", "--code", "rec"]
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="abides_ddqn_")
    os.makedirs(os.path.join(tmp, "data", "marketreplay", "level_1"))
    os.symlink(os.path.join(R.REF, "data", "lobster"), os.path.join(tmp, "data", "lobster"))
    os.chdir(tmp)
    sink = io.StringIO()
    try:
        import importlib
        with contextlib.redirect_stdout(sink):
            mod = importlib.import_module("config." + cfg)
    finally:
        os.chdir(cwd)
    tail = sink.getvalue().strip().splitlines()[-6:]
    print("\n".join(tail))

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

    agents = mod.agents
    replay = agents[1]
    od = replay.historical_orders.orders_dict
    stream = np.array([(R.REC.ns(ts), int(r["ORDER_ID"]), int(r["PRICE"]), int(r["SIZE"]), 1 if r["BUY_SELL_FLAG"] == "BUY" else 0)
                       for ts in od for r in od[ts]], dtype=np.int64)
    mom = [a for a in agents if type(a).__name__ == "MomentumAgent"]
    twap = [a for a in agents if type(a).__name__ in ("TWAPExecutionAgent", "VWAPExecutionAgent")][0]
    dq = agents[-1]
    assert type(dq).__name__ == "DDQLearningExecutionAgent"
    T = len(dq.experience)
    exp = np.full((T, 6), np.nan)
    for t, (s, a, sp, r) in dq.experience.items():
        exp[t, 0:2] = s
        exp[t, 2] = a
        exp[t, 3:5] = sp
        exp[t, 5] = np.nan if r is None else r
    obs = np.array([dq.observation[t] for t in range(T)], dtype=np.float64)
    hold = []
    for a in agents[1:]:
        hold.append((a.id, int(a.holdings.get(ticker, 0)), int(a.holdings["CASH"]), int(a.last_trade[ticker]) if ticker in a.last_trade else 0, len(a.orders)))
    keep = 60000
    data = dict(
        ticker=np.array(ticker), date=np.array(date), seed=np.array(seed), is_buy=np.array(int(direction == "BUY")), policy_seed=np.array(pseed), stream=stream,
        mom_sizes=np.array([a.size for a in mom], dtype=np.int64), n_pops=np.array(len(pops)), n_ops=np.array(len(ops)), n_notes=np.array(len(notes)),
        pop_hash_ckpt=np.array(ck, dtype=np.uint64), note_hash=np.array(hn, dtype=np.uint64), snap_hash=np.array(hs, dtype=np.uint64),
        pops_head=pops[:keep], ops_head=ops[: keep // 2], notes_head=notes[: keep // 2], snaps_head=snaps[: keep // 4],
        rl_ops=ops[ops[:, 2] >= 9],                     # every book op requested by the two execution agents (ids 9, 10)
        kind_counts=np.bincount(pops[:, 4], minlength=len(R.MSG_KINDS)),
        max_levels=np.array([snaps[:, 0].max(), snaps[:, 1].max()]), max_resting=np.array(snaps[:, 2].max()),
        holdings=np.array(hold, dtype=np.int64),
        actions=np.array(POLICY["actions"], dtype=np.int64), experience=exp, observation=obs,
        price_path=np.array(dq.price_path, dtype=np.float64), action_hist=np.array(dq.action_hist, dtype=np.float64),
        step_reward_hist=np.array(dq.step_reward_hist, dtype=np.float64),
        ddqn_final=np.array([dq.remaining_qty, dq.remaining_time, dq.t, dq.arrival_price, len(dq.executed_orders)], dtype=np.float64),
        twap_final=np.array([twap.rem_quantity, twap.arrival_price, len(twap.executed_orders)], dtype=np.float64),
    )
    if vwap:
        data["vwap_profile"], data["vwap_schedule"] = POLICY["vwap_profile"], POLICY["vwap_schedule"]
    if POLICY["mlp"] is not None:                       # the network that chose the actions, its Q rows, and a slimmer fixture (the traces are pinned by the other goldens)
        data["mlp_params"], data["mlp_q"] = POLICY["mlp"][0], np.array(POLICY["q"], dtype=np.float32)
        for k in ("pops_head", "ops_head", "notes_head", "snaps_head", "rl_ops"):
            data[k] = data[k][:2000]
        with np.load(os.path.join(os.path.dirname(HERE), "tests", "golden", "env_IBM_%s_s4242.npz" % date)) as ge:      # the same day is already a fixture
            assert np.array_equal(ge["stream"], data["stream"])
        data["stream_fixture"] = np.array("env_IBM_%s_s4242.npz" % date); del data["stream"]
    os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    np.savez_compressed(out, **data)
    print("recorded ddqn", ticker, date, "pops", len(pops), "ops", len(ops), "notes", len(notes), "ticks", T, "->", out)
    print("kind counts", dict(zip(R.MSG_KINDS, data["kind_counts"])))
    print("max levels", data["max_levels"], "max resting", data["max_resting"], "ddqn final", data["ddqn_final"], "twap final", data["twap_final"])
    print("sum reward", float(np.sum(dq.step_reward_hist)), "actions used", sorted(set(POLICY["actions"])))


if __name__ == "__main__":
    main()
