"""The reference's config-as-script surface: `python abides.py -c <config> <flags>` (abides.py:18-34 imports config/<config>.py, which parses its own
flags and calls Kernel.runner).  Here the same command line builds the abx_sim_config / environment of the batched simulator:

    cfg, run = configs.from_argv(["-c", "rmsc03", "-t", "ABM", "-d", "20190628", "-s", "1234", "--mm-pov", "0.1", "--mm-num-ticks", "10"])
    sim = run(n_envs=4096)          # BatchedSim, reset with seeds seed, seed + 1, ... (environment e seeded base + e)

Flags mirrored (same spelling, same defaults):
  sparse_zi_100 / sparse_zi_1000   -s/--seed, -l/--log_dir, -b/--book_freq, -o/--log_orders, -v (config/sparse_zi_1000.py:17-34; logging flags accepted, no effect)
  rmsc03                           -t/--ticker, -d/--historical-date, -s, -l, -v, --mm-pov, --mm-min-order-size, --mm-window-size, --mm-num-ticks,
                                   --mm-wake-up-freq, --wide-book (config/rmsc03.py:30-46)
  rmsc01 / rmsc02                  -s, -l, -v, -a/--agent_name (config/rmsc01.py:28-40; an --agent_name "agent under test" is a Python class and is rejected)
  marketreplay                     -t/--ticker, -d/--date, -l, -lvl/--level, -s, -v (config/marketreplay.py:19-28): ABIDESEnv with order_level 0
"""
import argparse
import re

import numpy as np

from .env import ABIDESEnv, env_config, lobster_message_path, load_lobster_csv
from .sim import BatchedSim, rmsc01_config, rmsc02_config, rmsc03_config, sparse_zi_config

NS = 10 ** 9
_UNITS = {"ns": 1, "us": 10 ** 3, "ms": 10 ** 6, "s": NS, "S": NS, "sec": NS, "min": 60 * NS, "T": 60 * NS, "h": 3600 * NS, "H": 3600 * NS}


def timedelta_ns(text):
    """pd.Timedelta(text) for the '<number><unit>' strings the configs pass ("1S", "30s", "1min"): nanoseconds."""
    m = re.fullmatch(r"\s*([0-9.]+)\s*([A-Za-z]+)\s*", str(text))
    if not m or m.group(2) not in _UNITS:
        raise ValueError("cannot parse time delta %r" % (text,))
    return int(round(float(m.group(1)) * _UNITS[m.group(2)]))


def _base_parser(config):
    p = argparse.ArgumentParser(prog="abides.py -c " + config, add_help=False)
    p.add_argument("-c", "--config", default=config)
    p.add_argument("-l", "--log_dir", default=None)
    p.add_argument("-s", "--seed", type=int, default=None)
    p.add_argument("-v", "--verbose", action="store_true")
    p.add_argument("--config_help", action="store_true")
    return p


def from_argv(argv, lib=None):
    """(config object, run(n_envs, device=0, **overrides) -> ready simulator) for an `abides.py` command line.  `-s` absent: the reference seeds from the
    clock (config/sparse_zi_1000.py:69-71); here the base seed is then 0 and says so in run.seed."""
    argv = list(argv)
    if "-c" not in argv and "--config" not in argv:
        raise SystemExit("abides.py: -c/--config is required")
    config = argv[argv.index("-c" if "-c" in argv else "--config") + 1]
    p = _base_parser(config)
    if config in ("sparse_zi_100", "sparse_zi_1000"):
        p.add_argument("-b", "--book_freq", default=None)
        p.add_argument("-o", "--log_orders", action="store_true")
        a, _ = p.parse_known_args(argv)
        cfg = sparse_zi_config(100 if config.endswith("100") else 1000, lib=lib)
    elif config == "rmsc03":
        p.add_argument("-t", "--ticker", required=True)
        p.add_argument("-d", "--historical-date", required=True)
        p.add_argument("--mm-pov", type=float, default=0.05)
        p.add_argument("--mm-min-order-size", type=int, default=20)
        p.add_argument("--mm-window-size", type=int, default=5)
        p.add_argument("--mm-num-ticks", type=int, default=20)
        p.add_argument("--mm-wake-up-freq", type=str, default="1S")
        p.add_argument("--wide-book", action="store_true")
        a, _ = p.parse_known_args(argv)
        cfg = rmsc03_config(lib=lib, mm_pov=a.mm_pov, mm_min_order_size=a.mm_min_order_size, mm_window_size=a.mm_window_size, mm_num_ticks=a.mm_num_ticks,
                            mm_wake_ns=timedelta_ns(a.mm_wake_up_freq))
    elif config in ("rmsc01", "rmsc02"):
        p.add_argument("-a", "--agent_name", default=None)
        a, _ = p.parse_known_args(argv)
        if a.agent_name is not None:
            raise SystemExit("abides.py -c %s -a %s: a user-defined agent under test is a Python class; outside the batched simulator's scope" % (config, a.agent_name))
        cfg = (rmsc01_config if config == "rmsc01" else rmsc02_config)(lib=lib)
    elif config == "marketreplay":
        p.add_argument("-t", "--ticker", required=True)
        p.add_argument("-d", "--date", required=True)
        p.add_argument("-lvl", "--level", default="1")
        a, _ = p.parse_known_args(argv)
        cfg = env_config(lib=lib, order_level=0, stop_ns=(16 * 3600 + 60) * NS, queue_cap=256, level_cap=1024)     # config/marketreplay.py:60-62 kernel stop 16:01
    else:
        raise SystemExit("abides.py -c %s: this config is outside the batched simulator's scope (sparse_zi_100, sparse_zi_1000, rmsc01, rmsc02, rmsc03, marketreplay)" % config)
    seed = 0 if a.seed is None else int(a.seed)

    def run(n_envs=1, device=0, data_root="data/lobster", lib_path=None, **overrides):
        for k, v in overrides.items():
            if not hasattr(cfg, k):
                raise AttributeError("config has no field %r" % k)
            setattr(cfg, k, v)
        if config == "marketreplay":
            stream = load_lobster_csv(lobster_message_path(a.ticker, a.date, data_root, int(a.level), dated_folder=True))
            env = ABIDESEnv(stream, n_envs=n_envs, device=device, cfg=cfg, lib_path=lib_path)
            env.reset()
            return env
        sim = BatchedSim(cfg, n_envs, device=device, lib_path=lib_path)
        if cfg.rng_mode == 0:
            sim.reset(np.arange(n_envs, dtype=np.uint64) + np.uint64(seed))
        return sim
    run.seed, run.args, run.config = seed, a, config
    return cfg, run
