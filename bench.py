#!/usr/bin/env python3
"""bench.py -- LOB messages/s of the batched ABIDES simulator on the sparse_zi_1000 shape.

  python bench.py --gpus N --steps K --warmup W            our arm (hand-written sm_100a kernels through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  reference arm: the reference's CPU algorithm (oracle port,
                                                            the Python reference cannot travel to the GPU box) on all host cores

Workload (BASELINE.json configs[1]): config/sparse_zi_1000.py -- 1000 ZeroIntelligenceAgents + exchange, sparse OU
oracle -- 16 384 independent environments per GPU, environment e seeded base+e, GPU-native Philox streams.
A "step" advances every environment's event loop (Kernel.py:190-292) by `--slice-s` simulated seconds with ONE launch
of abx_run_kernel.  A "message" is one event-queue pop counted exactly like the reference's ttl_messages
(Kernel.py:211).  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NS = 10 ** 9
FIXTURE = os.path.join(ROOT, "tests", "golden", "env_IBM_2003-01-14_s789.npz")   # LOBSTER sample day as the reference parsed it
DQ_FIXTURE = os.path.join(ROOT, "tests", "golden", "ddqn_IBM_2003-01-14_s4242.npz")   # same day, with the recorded MomentumAgent sizes
QNET_FLOP_PER_ROW = 2 * (2 * 32 + 32 * 64 + 64 * 128 + 128 * 128 + 128 * 64 + 64 * 32 + 32 * 24)   # util/model/QNets.py:7-27 with the 2-entry state
DAY_FIXTURES = ("env_IBM_2003-01-14_s789.npz", "env_IBM_2003-01-15_s4242.npz", "ddqn_IBM_2003-01-16_s99_sell.npz")   # three IBM LOBSTER sample days as the reference parsed them


def replay_days():
    """The replayed order streams, round robin over environments (IBM sample days; environment e replays day e % 3)."""
    import numpy as np
    out = []
    for f in DAY_FIXTURES:
        with np.load(os.path.join(ROOT, "tests", "golden", f)) as g:
            out.append(g["stream"].copy())
    return out


B_MSG = 320                 # algorithmic bytes per LOB message (DESIGN.md "Roofline")
PUBLISHED_MSGS_PER_S = 3100.4   # reference's own sparse_zi_1000 run (tests/sparse_zi_1000.txt:22)


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL's version banner) write to fd 1; the contract is ONE JSON line on stdout, so everything else goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit_json(obj):
    f = _REAL_STDOUT or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_peak_tflops():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["bf16_tflops"]), "measured burst cuBLAS bf16 (MEASURED_PEAKS.json)"
    except Exception:
        return 1590.0, "fallback (B200_PROFILING.md)"


def shared_config(args):
    """The workload both arms are measured on (BASELINE.json configs[1]); arm-specific facts go under the line's "detail" key, so that the two
    arms' `config` objects are identical."""
    return {"workload": "config/sparse_zi_%d.py: %d ZeroIntelligenceAgents + ExchangeAgent, sparse mean-reverting OU oracle with megashocks, 00:00-17:00 (market 09:30-16:00); "
                        "independent environments, environment e seeded %d+e; one message = one pop of the kernel event queue, counted like ttl_messages (Kernel.py:211)"
                        % (args.variant, args.variant, args.seed),
            "envs_per_gpu": args.envs_per_gpu, "variant": args.variant, "seed": args.seed}


def python_reference_record():
    """The unmodified Python reference under the config/parallel.py pattern, measured in the BUILD CONTAINER by tools/measure_python_reference.py (the
    reference tree cannot travel to the GPU box, and nothing here may read it at run time): reported with its provenance, not as a same-box number."""
    try:
        r = json.load(open(os.path.join(ROOT, "profiles", "r02_python_reference_cpu.json")))
        return {"value": r["msgs_per_s_event_loop"], "unit": "msgs/s", "processes": r["processes"], "cores": r["cores"], "where": r["where"], "what": r["what"],
                "msgs_per_s_wall": r["msgs_per_s_wall"], "source": "profiles/r02_python_reference_cpu.json (tools/measure_python_reference.py)"}
    except Exception:
        return None


def sim_closed(sim):
    return getattr(sim, "_h", None) is None or not sim._h


def profiled_traffic():
    """dram bytes per launch of abx_run_kernel from the committed ncu --set full capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "run_kernel_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def oracle_msgs_per_s(n_env_days, threads, variant=1000, seed0=5000):
    """Reference CPU algorithm (oracle/abides_oracle.c, scalar C port) on `threads` host threads: full env-days."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import OracleSim, lib
    lib()

    def one(i):
        s = OracleSim(variant, seed0 + i, 0)
        s.start()
        t0 = time.perf_counter()
        n, _ = s.run_until(2 ** 62)          # event loop only, like Kernel.py:184,301 (construction excluded)
        return n, time.perf_counter() - t0

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:       # ctypes releases the GIL during the C call
        res = list(ex.map(one, range(n_env_days)))
    wall = time.perf_counter() - t0
    msgs = sum(r[0] for r in res)
    return msgs, wall, sum(r[1] for r in res)


def oracle_env_steps_per_s(n_episodes, steps, threads):
    """Reference CPU algorithm of ABIDESEnv.step (oracle port) on `threads` host threads: `steps` steps after the first."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import OracleEnv, lib
    lib()
    days = replay_days()

    def one(i):
        rng = np.random.RandomState(100 + i)
        env = OracleEnv(days[i % len(days)])
        env.step(np.array([0.02, 0.5, 0.5]))              # 00:00 -> 09:40 start-up, untimed like the GPU arm
        n0 = env.n_pops
        t0 = time.perf_counter()
        for _ in range(steps):
            env.step(np.array([rng.uniform(0, 0.04), rng.uniform(), rng.uniform()]))
        return steps, env.n_pops - n0, time.perf_counter() - t0

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(one, range(n_episodes)))
    wall = time.perf_counter() - t0 - 0.0
    busy = max(r[2] for r in res)
    return sum(r[0] for r in res), sum(r[1] for r in res), sum(r[2] for r in res) / threads if n_episodes >= threads else busy, wall


def oracle_ddqn_ticks_per_s(n_episodes, ticks, threads):
    """Reference CPU algorithm of the DDQN execution config (oracle port; the Q-network is replaced by random actions) on host threads."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import OracleDDQNEnv, lib
    lib()
    with np.load(DQ_FIXTURE) as g:                         # materialise before the threads start: NpzFile is not thread safe
        dq_sizes = g["mom_sizes"].copy()
    days = replay_days()

    def one(i):
        rng = np.random.RandomState(300 + i)
        env = OracleDDQNEnv(days[i % len(days)], dq_sizes)
        env.step(0)                                        # 00:00 -> 10:00 start-up, untimed like the GPU arm
        n0 = env.n_pops
        t0 = time.perf_counter()
        for _ in range(ticks):
            env.step(int(rng.randint(0, 24)))
        return ticks, env.n_pops - n0, time.perf_counter() - t0

    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(one, range(n_episodes)))
    busy = sum(r[2] for r in res) / threads if n_episodes >= threads else max(r[2] for r in res)
    return sum(r[0] for r in res), sum(r[1] for r in res), busy


def reference_ddqn_block(cores):
    st, msgs, busy = oracle_ddqn_ticks_per_s(max(4 * cores, 8), 400, cores)
    return {"metric": "DDQN execution env ticks/sec", "value": st / busy, "unit": "steps/s", "msgs_per_s": msgs / busy, "cores": cores, "kind": "port",
            "sample": "%d runs x 400 decision ticks after the 10:00 start-up (IBM 2003-01-14/15/16 round robin, random actions, no network), %d threads" % (max(4 * cores, 8), cores)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(4 * cores, 16)
    oracle_msgs_per_s(min(cores, 8), cores, args.variant)                       # warm-up (library load, page-in)
    for _ in range(max(args.warmup - 1, 0)):
        oracle_msgs_per_s(per_step, cores, args.variant)
    tot_m, tot_w = 0, 0.0
    for k in range(args.steps):
        m, w, _ = oracle_msgs_per_s(per_step, cores, args.variant, seed0=9000 + k * per_step)
        tot_m += m; tot_w += w
    v = tot_m / tot_w
    sample = "%d steps x %d full env-days of sparse_zi_%d (event loop only), %d threads" % (args.steps, per_step, args.variant, cores)
    emit_json({
        "impl": "reference", "metric": "LOB msgs/sec", "value": v, "unit": "msgs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_w / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": v / PUBLISHED_MSGS_PER_S, "dtype": "int64+f64", "data": "synthetic",
        "config": shared_config(args),
        "detail": {"arm": "reference CPU algorithm: oracle/abides_oracle.c (C port of the Python reference, pinned to it bit for bit) on all host threads; whole environment-days from MT19937 "
                          "seeds, event loop only (construction excluded like Kernel.py:184,301)", "sample": sample, "python_reference": python_reference_record()},
        "cpu_baseline": {"value": v, "unit": "msgs/s", "cores": cores, "kind": "port", "sample": sample, "python_reference": python_reference_record()},
        "e2e": {"value": v, "unit": "msgs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "rmsc03": reference_rmsc03_block(cores), "env": reference_env_block(cores), "ddqn": reference_ddqn_block(cores),
    })


def reference_env_block(cores):
    st, msgs, busy, wall = oracle_env_steps_per_s(max(2 * cores, 4), 400, cores)
    return {"metric": "ABIDESEnv steps/sec", "value": st / busy, "unit": "steps/s", "msgs_per_s": msgs / busy, "cores": cores, "kind": "port",
            "sample": "%d episodes x 400 steps after the 09:40 start-up (IBM 2003-01-14/15/16, round robin), %d threads" % (max(2 * cores, 4), cores)}


def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    from marl_optimal_execution_b200 import _lib, distributed as D
    from marl_optimal_execution_b200.sim import BatchedSim, sparse_zi_config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = sparse_zi_config(args.variant)
    n_envs = args.envs_per_gpu
    sim = BatchedSim(cfg, n_envs, device=local_rank)
    lo = rank * n_envs
    seeds = D.env_seeds(args.seed, lo, lo + n_envs)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    # simulated-time budget: prefix + (W + 2K) slices must stay inside market hours
    t_open, t_close = int(cfg.mkt_open_ns), int(cfg.mkt_close_ns)
    prefix = 120 * NS
    n_slices = args.warmup + 2 * args.steps
    slice_ns = min(int(args.slice_s * NS), (t_close - t_open - prefix - 60 * NS) // max(n_slices, 1))
    sim.reset(seeds, stream=sp)
    t = t_open + prefix
    sim.run(t, stream=sp)                               # 00:00 -> 09:32 start-up transient, untimed
    for _ in range(args.warmup):
        t += slice_ns
        sim.run(t, stream=sp)
    torch.cuda.synchronize(dev)
    st0 = sim.stats(stream=sp)
    launches0 = sim.launch_count

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    D.barrier(); torch.cuda.synchronize(dev)
    ev[0].record(stream)
    for k in range(args.steps):                         # ---- timed region: K launches of abx_run_kernel, state resident in HBM
        t += slice_ns
        sim.run(t, stream=sp)
        ev[k + 1].record(stream)
    torch.cuda.synchronize(dev); D.barrier()
    launches = sim.launch_count - launches0
    elapsed_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    clocks = sampler.stop() if sampler else None
    st1 = sim.stats(stream=sp)
    msgs_local = int(st1["messages"].sum() - st0["messages"].sum())
    err_envs = int(((st1["flags"] & _lib.F_ERROR_MASK) != 0).sum())

    # ---- e2e: same metric through the C ABI with HOST buffers (pinned): per step H2D of the per-env horizons,
    #      launch, D2H of every environment's abx_env_stats record; wall clock around sync points.  SAME simulated interval as the
    #      device-timed loop: the batch is reset with the same seeds and brought to the same point (untimed) first.
    until_pin = torch.empty(n_envs, dtype=torch.int64).pin_memory()
    stats_pin = torch.empty(n_envs * ctypes.sizeof(_lib.EnvStats), dtype=torch.uint8).pin_memory()
    stats_np = stats_pin.numpy().view(_lib.STATS_DTYPE)
    sim.reset(seeds, stream=sp)
    t = t_open + prefix
    sim.run(t, stream=sp)
    for _ in range(args.warmup):
        t += slice_ns
        sim.run(t, stream=sp)
    torch.cuda.synchronize(dev)
    st2 = sim.stats(stream=sp)
    assert int(st2["messages"].sum()) == int(st0["messages"].sum()), "the e2e leg does not start where the device-timed leg started"
    D.barrier(); torch.cuda.synchronize(dev)
    w0 = time.perf_counter()
    for k in range(args.steps):
        t += slice_ns
        until_pin.fill_(t)
        sim.run_each(until_pin.data_ptr(), stream=sp)
        sim.stats(stream=sp, out=stats_np)              # synchronises the stream
    torch.cuda.synchronize(dev)
    e2e_s_local = time.perf_counter() - w0
    D.barrier()
    e2e_msgs_local = int(stats_np["messages"].sum() - st2["messages"].sum())
    assert e2e_msgs_local == msgs_local, "e2e and device-timed legs must cover the same messages"

    # ---- whole environment-days (what the reference arm times): fresh seeds, 00:00 -> 17:00 in ONE launch per batch, start-up transient included
    wd_local = None
    if not args.no_whole_day:
        sim.reset(D.env_seeds(args.seed + 7777777, lo, lo + n_envs), stream=sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier(); torch.cuda.synchronize(dev)
        l_wd = sim.launch_count
        e0.record(stream); sim.run(stream=sp); e1.record(stream)
        torch.cuda.synchronize(dev); D.barrier()
        stw = sim.stats(stream=sp)
        wd_local = {"msgs": int(stw["messages"].sum()), "ms": e0.elapsed_time(e1), "errs": int(((stw["flags"] & _lib.F_ERROR_MASK) != 0).sum()), "launches": sim.launch_count - l_wd}

    # ---- second headline metric: ABIDESEnv steps/s (Exchange + MarketReplayAgent + RL execution agent under GymKernel)
    env_local = None
    state_bytes = sim.device_bytes
    r3_local = None
    if not args.no_rmsc03:
        sim.close()
        r3_local = bench_rmsc03(args, rank, local_rank, dev, stream, sp)
    r1_local = r2_local = None
    if not args.no_rmsc01:
        if not sim_closed(sim):
            sim.close()
        r1_local = bench_rmsc01(args, rank, local_rank, dev, stream, sp)
        r2_local = bench_rmsc01(args, rank, local_rank, dev, stream, sp, subscriptions=True)
    mr_local = None
    if not args.no_marketreplay:
        if not sim_closed(sim):
            sim.close()
        mr_local = bench_marketreplay(args, rank, local_rank, dev, stream, sp)
    if not args.no_env:
        if not sim_closed(sim):
            sim.close()
        env_local = bench_env(args, rank, local_rank, dev, stream, sp)

    dq_local = None
    if not args.no_ddqn:
        if not sim_closed(sim):
            sim.close()
        dq_local = bench_ddqn(args, rank, local_rank, dev, stream, sp)

    # ---- aggregate over ranks: sums by all-gather of the summary vectors (NCCL), times by max
    g = D.gather_summaries(torch.tensor([msgs_local, e2e_msgs_local, err_envs], dtype=torch.int64), device=dev)
    elapsed_ms = D.max_over_ranks(elapsed_ms, device=dev)
    e2e_s = D.max_over_ranks(e2e_s_local, device=dev)
    r3_block = None
    if r3_local is not None:
        r3_block = {"metric": "LOB msgs/sec (rmsc03)", "unit": "msgs/s", "envs_per_gpu": args.rmsc03_envs_per_gpu,
                    "workload": "config/rmsc03.py shape: 50 NoiseAgents + 10 ValueAgents + 2 MomentumAgents + POVMarketMakerAgent + exchange, 09:30-09:46, "
                                "zero latency, %d envs/GPU (Philox streams); the whole run of every environment in one abx_run_kernel launch; "
                                "2 timed runs on fresh seeds after one warm-up run, resets untimed" % args.rmsc03_envs_per_gpu}
        for key in ("plain", "pov"):
            r = r3_local[key]
            gr = D.gather_summaries(torch.tensor([r["msgs"], r["errs"], r["invalid"], r["runs"]], dtype=torch.int64), device=dev)
            t_r = D.max_over_ranks(r["ms"], device=dev) / 1e3
            blk = {"value": int(gr[:, 0].sum()) / t_r, "messages_per_env_run": int(gr[:, 0].sum()) / max(int(gr[:, 3].sum()), 1),
                   "ms_per_run": 1e3 * t_r / 2, "error_envs": int(gr[:, 1].sum()), "gpu_launches": int(r["launches"])}
            if key == "plain":
                r3_block.update(blk)
                ach = int(gr[:, 0].sum()) / world * B_MSG / t_r / 1e9
                r3_block["roofline"] = {"bound": "hbm", "achieved": ach, "peak": measured_peak_hbm()[0], "unit": "GB/s", "frac": ach / measured_peak_hbm()[0],
                                        "traffic": None, "kernel": "abx_run_kernel<0,2,0,2>", "algorithmic_bytes_per_msg": B_MSG,
                                        "note": "profiles/r01_rmsc03_run_kernel_ncu_full_summary.txt: the state of 4 096 environments sits in the L2; latency bound"}
            else:
                blk["one_sided_book_envs"] = int(gr[:, 2].sum())
                blk["note"] = ("same population + POVExecutionAgent (agent/execution/baselines/pov_agent.py; BUY 120 000 at 50 % of volume, 09:32-09:43); "
                               "one_sided_book_envs = runs in which the agent saw an empty book side (the reference raises TypeError there; flagged F_OBS_INVALID, the run continues)")
                r3_block["with_pov_execution_agent"] = blk
    r1_block = None
    if r1_local is not None:
        g1 = D.gather_summaries(torch.tensor([r1_local["msgs"], r1_local["errs"], r1_local["hbl_orders"]], dtype=torch.int64), device=dev)
        t1 = D.max_over_ranks(r1_local["ms"], device=dev) / 1e3
        r1_block = {"metric": "LOB msgs/sec (rmsc01)", "unit": "msgs/s", "value": int(g1[:, 0].sum()) / t1, "envs_per_gpu": args.rmsc01_envs_per_gpu, "ms_per_run": 1e3 * t1,
                    "messages_per_env_run": int(g1[:, 0].sum()) / (world * args.rmsc01_envs_per_gpu), "error_envs": int(g1[:, 1].sum()), "gpu_launches": int(r1_local["launches"]),
                    "limit_orders_per_env_run": int(g1[:, 2].sum()) / (world * args.rmsc01_envs_per_gpu),
                    "workload": "config/rmsc01.py shape: MarketMakerAgent + 50 ZI + 25 HeuristicBeliefLearningAgents (QUERY_ORDER_STREAM, L = 2) + 24 MomentumAgents + exchange, "
                                "09:30-10:00 of the day, zero latency, %d envs/GPU (Philox streams), one abx_run_kernel<0,2,0,5> launch; one timed run on fresh seeds after a "
                                "warm-up run, resets untimed" % args.rmsc01_envs_per_gpu}
    if r2_local is not None:
        g2 = D.gather_summaries(torch.tensor([r2_local["msgs"], r2_local["errs"], r2_local["hbl_orders"]], dtype=torch.int64), device=dev)
        t2 = D.max_over_ranks(r2_local["ms"], device=dev) / 1e3
        r1_block["rmsc02"] = {"metric": "LOB msgs/sec (rmsc02)", "unit": "msgs/s", "value": int(g2[:, 0].sum()) / t2, "ms_per_run": 1e3 * t2,
                              "messages_per_env_run": int(g2[:, 0].sum()) / (world * args.rmsc01_envs_per_gpu), "error_envs": int(g2[:, 1].sum()), "gpu_launches": int(r2_local["launches"]),
                              "workload": "config/rmsc02.py shape: the rmsc01 population with MARKET_DATA subscriptions (market maker + 24 momentum agents), pairwise latency U(21 us, 13 ms) + "
                                          "6-entry noise; the WHOLE day 00:00-17:00 of %d envs/GPU in one abx_run_kernel<0,0,0,5> launch" % args.rmsc01_envs_per_gpu}
    env_block = None
    if env_local is not None:
        ge = D.gather_summaries(torch.tensor([env_local["steps"], env_local["msgs"], env_local["e2e_steps"], env_local["errs"]], dtype=torch.int64), device=dev)
        t_env = D.max_over_ranks(env_local["ms"], device=dev) / 1e3
        t_e2e = D.max_over_ranks(env_local["e2e_s"], device=dev)
        env_block = {"metric": "ABIDESEnv steps/sec", "value": int(ge[:, 0].sum()) / t_env, "unit": "steps/s", "msgs_per_s": int(ge[:, 1].sum()) / t_env,
                     "envs_per_gpu": args.env_envs_per_gpu, "steps": args.env_steps, "ms_per_step": 1e3 * t_env / args.env_steps,
                     "messages_per_env_step": int(ge[:, 1].sum()) / max(int(ge[:, 0].sum()), 1), "error_envs": int(ge[:, 3].sum()),
                     "e2e": {"value": int(ge[:, 2].sum()) / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": 24 * args.env_envs_per_gpu,
                             "d2h_bytes_per_step": 81 * args.env_envs_per_gpu},
                     "gpu_launches": env_local["launches"], "workload": "ABIDESEnv.py shape: exchange + MarketReplayAgent (IBM 2003-01-14/15/16 LOBSTER sample days, round robin over environments) + "
                     "DummyRLExecutionAgent (BUY 1e5, 30 s, order_level 2), random actions; one abx_env_step_kernel launch per step; the timed window is the whole episode "
                     "(761 ticks) minus the start-up and warm-up steps"}
    dq_block = None
    if dq_local is not None:
        gd = D.gather_summaries(torch.tensor([dq_local["steps"], dq_local["msgs"], dq_local["e2e_steps"], dq_local["errs"], dq_local["train_done"]], dtype=torch.int64), device=dev)
        t_dq = D.max_over_ranks(dq_local["ms"], device=dev) / 1e3
        t_dq_e2e = D.max_over_ranks(dq_local["e2e_s"], device=dev)
        t_dq_train = D.max_over_ranks(dq_local["train_s"], device=dev)
        nd, Kd = args.ddqn_envs_per_gpu, args.ddqn_steps
        tf_peak = measured_peak_tflops()
        q_ach = nd * QNET_FLOP_PER_ROW / (dq_local["qnet_ms"] / 1e3) / 1e12
        dq_block = {"metric": "DDQN execution env ticks/sec (Q-network forward + environment step)", "value": int(gd[:, 0].sum()) / t_dq, "unit": "steps/s",
                    "msgs_per_s": int(gd[:, 1].sum()) / t_dq, "envs_per_gpu": nd, "steps": Kd, "ms_per_step": 1e3 * t_dq / Kd,
                    "messages_per_env_step": int(gd[:, 1].sum()) / max(int(gd[:, 0].sum()), 1), "error_envs": int(gd[:, 3].sum()),
                    "e2e": {"value": int(gd[:, 2].sum()) / t_dq_e2e, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 121 * nd,
                            "note": "actions are produced on the device by the Q-network; every tick obs, experience tuple, reward and done are read back to pinned host memory"},
                    "training": {"value": int(gd[:, 0].sum()) / t_dq_train, "unit": "steps/s", "learn_steps": dq_local["learn_steps"], "batch": args.ddqn_batch,
                                 "replay_buffer_rows": dq_local["buffer"], "finished_envs": int(gd[:, 4].sum()), "note": "same loop with the learner on: experience tuples into a device replay buffer, one "
                                 "train_neural_nets-style update (PyTorch fp32 autograd, RMSprop) every 5 ticks, new weights packed into the tcgen05 operand image on the device; no host synchronisation inside the loop; wall clock; finished_envs = episodes whose parent order the learned policy completed before the last training tick (they idle from then on); "
                                 "with N ranks one policy is trained: the 38 k gradients are averaged by an NCCL all-reduce before every update"},
                    "gpu_launches": dq_local["launches"], "dtype": "int64+f64 (environment), bf16x3 -> fp32 accumulate (Q-network)",
                    "qnet_roofline": {"bound": "tensor", "achieved": q_ach, "peak": tf_peak[0], "unit": "TFLOP/s", "frac": q_ach / tf_peak[0], "traffic": None,
                                      "peak_source": tf_peak[1], "kernel": "abx_qnet_forward_kernel", "kernel_ms": dq_local["qnet_ms"],
                                      "algorithmic_flop_per_row": QNET_FLOP_PER_ROW, "share_of_tick": dq_local["qnet_ms"] * Kd / dq_local["ms"],
                                      "note": "7 dependent layers per 128-row tile, %d tiles on 148 SMs: latency bound by construction, not a throughput kernel" % ((nd + 127) // 128)},
                    "workload": "config/execution/marketreplay/execution_marketreplay_ddqn.py shape: exchange + MarketReplayAgent (IBM 2003-01-14/15/16 LOBSTER sample days, round robin) + 7 MomentumAgents + "
                    "TWAPExecutionAgent + DDQLearningExecutionAgent (BUY 5e5, 30 s ticks from 10:00), epsilon-greedy (0.9) actions from a random-init 2-32-64-128-128-64-32-24 network; "
                    "per tick one abx_qnet_forward_kernel + one abx_dq_step_kernel launch"}
    wd_block = None
    if wd_local is not None:
        gw = D.gather_summaries(torch.tensor([wd_local["msgs"], wd_local["errs"]], dtype=torch.int64), device=dev)
        t_wd = D.max_over_ranks(wd_local["ms"], device=dev) / 1e3
        wd_block = {"metric": "LOB msgs/sec, whole environment-days", "value": int(gw[:, 0].sum()) / t_wd, "unit": "msgs/s", "ms": 1e3 * t_wd, "messages_per_env_day": int(gw[:, 0].sum()) / (n_envs * world),
                    "error_envs": int(gw[:, 1].sum()), "gpu_launches": wd_local["launches"],
                    "note": "the interval the reference arm times: %d environments per GPU from fresh seeds, 00:00 -> 17:00 (start-up wake-ups, open, close, after-hours) in one abx_run_kernel launch" % n_envs}
    mr_block = None
    if mr_local is not None:
        gm = D.gather_summaries(torch.tensor([mr_local["msgs"], mr_local["errs"]], dtype=torch.int64), device=dev)
        t_mr = D.max_over_ranks(mr_local["ms"], device=dev) / 1e3
        mr_block = {"metric": "LOB msgs/sec (marketreplay)", "value": int(gm[:, 0].sum()) / t_mr, "unit": "msgs/s", "envs_per_gpu": args.mr_envs_per_gpu, "ms": 1e3 * t_mr,
                    "messages_per_env_day": int(gm[:, 0].sum()) / (args.mr_envs_per_gpu * world), "error_envs": int(gm[:, 1].sum()), "gpu_launches": mr_local["launches"],
                    "workload": "config/marketreplay.py shape (BASELINE configs[4]): ExchangeAgent + MarketReplayAgent replaying the GOOG 2012-06-21 LOBSTER sample day (49 482 rows, "
                                "the committed fixture the reference parsed) through the GPU books, one whole day per environment in one abx_env_step_kernel launch"}
    if rank != 0:
        return
    msgs, e2e_msgs, errs = int(g[:, 0].sum()), int(g[:, 1].sum()), int(g[:, 2].sum())
    value = msgs / (elapsed_ms / 1e3)
    peak, peak_src = measured_peak_hbm()
    # roofline of the dominant (only) kernel in the timed region, per launch, on rank 0's launches
    kern_ms = statistics.mean(step_ms)
    alg_bytes_per_launch = (msgs_local / args.steps) * B_MSG
    achieved = alg_bytes_per_launch / (kern_ms / 1e3) / 1e9
    traffic = profiled_traffic()
    out = {
        "metric": "LOB msgs/sec", "value": value, "unit": "msgs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": value / PUBLISHED_MSGS_PER_S, "dtype": "int64+f64", "data": "synthetic",
        "config": shared_config(args),
        "detail": {"arm": "hand-written sm_100a kernels through the C ABI, GPU-native Philox streams", "step": "%.1f simulated seconds per environment per launch (mid-day slices after a "
                   "09:30+120 s start-up transient)" % (slice_ns / NS), "messages_per_step": msgs // args.steps, "state_bytes_per_gpu": state_bytes,
                   "l2_policy": "inputs larger than L2: %.1f GB of per-environment state streamed per step, no flush needed" % (state_bytes / 1e9),
                   "vs_baseline_ref": "BASELINE.md section 1: 3100.4 msgs/s, the reference's own published single-process figure (i7 2.6 GHz); not a same-box measurement",
                   "error_envs": errs},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic["dram_bytes_per_message"] * (msgs_local / args.steps) if traffic else None,
                     "traffic_source": traffic["source"] if traffic else None, "peak_source": peak_src,
                     "kernel": "abx_run_kernel<PHILOX, MATRIX_NOISE, no instrumentation>", "kernel_ms": kern_ms, "algorithmic_bytes_per_msg": B_MSG,
                     "limiter": "instruction issue / instruction fetch of one dependent event chain per environment, not memory bandwidth (profiles/)",
                     "warp_instructions_per_message": traffic.get("warp_instructions_per_message") if traffic else None,
                     "issue_frac": ((msgs_local / args.steps) / (kern_ms / 1e3) * traffic["warp_instructions_per_message"] / (148 * 4 * (clocks["sm_mhz"] or 1965.0) * 1e6)
                                    if traffic and traffic.get("warp_instructions_per_message") and clocks else None),
                     "issue_frac_note": "warp-instructions issued per second / (148 SMs x 4 schedulers x SM clock); instructions per message from the committed ncu capture",
                     "note": "latency-bound by design (one dependent event chain per environment); see DESIGN.md"},
        "e2e": {"value": e2e_msgs / e2e_s, "unit": "msgs/s", "h2d_bytes_per_step": 8 * n_envs * 1,
                "d2h_bytes_per_step": ctypes.sizeof(_lib.EnvStats) * n_envs},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if wd_block is not None:
        out["whole_day"] = wd_block
    if mr_block is not None:
        out["marketreplay"] = mr_block
    if r3_block is not None:
        out["rmsc03"] = r3_block
    if r1_block is not None:
        out["rmsc01"] = r1_block
    if not args.no_env:
        out["env"] = env_block
    if dq_block is not None:
        out["ddqn"] = dq_block
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_days = max(12 * cores, 64)                          # ~0.07 CPU-s per env-day: 10-20 s of CPU work
        oracle_msgs_per_s(min(cores, 4), cores, args.variant)
        m, w, cpu_s = oracle_msgs_per_s(n_days, cores, args.variant)
        out["cpu_baseline"] = {"value": m / w, "unit": "msgs/s", "cores": cores, "kind": "port",
                               "sample": "%d full env-days of sparse_zi_%d, event loop only, %d threads, %.1f CPU-s" % (n_days, args.variant, cores, cpu_s),
                               "single_thread_value": m / cpu_s, "python_reference": python_reference_record()}
        if not args.no_rmsc03:
            out["cpu_baseline"]["rmsc03"] = reference_rmsc03_block(cores)
        if not args.no_env:
            out["cpu_baseline"]["env"] = reference_env_block(cores)
        if not args.no_ddqn:
            out["cpu_baseline"]["ddqn"] = reference_ddqn_block(cores)
    emit_json(out)


def bench_rmsc03(args, rank, local_rank, dev, stream, sp):
    """BASELINE configs[2]: config/rmsc03.py population (50 noise + 10 value + 2 momentum agents + POV market maker), without and with the POV
    execution agent; one abx_run_kernel launch runs the whole 09:30-09:46 simulation of every environment.  Reset (construction) is untimed,
    like Kernel.py:184,301; one warm-up run, then `reps` timed runs on fresh seeds."""
    import torch
    from marl_optimal_execution_b200 import _lib, distributed as D
    from marl_optimal_execution_b200.sim import BatchedSim, rmsc03_config
    n, reps, out = args.rmsc03_envs_per_gpu, 2, {}
    for pov in (False, True):
        sim = BatchedSim(rmsc03_config(pov_exec=pov), n, device=local_rank)
        ms, msgs, errs, invalid, l0 = 0.0, 0, 0, 0, 0
        for rep in range(reps + 1):
            sim.reset(D.env_seeds(args.seed + 100003 * rep, rank * n, (rank + 1) * n), stream=sp)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            D.barrier(); torch.cuda.synchronize(dev)
            lc = sim.launch_count
            e0.record(stream)
            sim.run(stream=sp)
            e1.record(stream)
            torch.cuda.synchronize(dev); D.barrier()
            if rep > 0:
                l0 += sim.launch_count - lc               # launches inside the timed regions only
                st = sim.stats(stream=sp)
                ms += e0.elapsed_time(e1); msgs += int(st["messages"].sum())
                errs += int(((st["flags"] & (_lib.F_ERROR_MASK & ~_lib.F_OBS_INVALID)) != 0).sum())
                invalid += int(((st["flags"] & _lib.F_OBS_INVALID) != 0).sum())
        out["pov" if pov else "plain"] = {"msgs": msgs, "ms": ms, "errs": errs, "invalid": invalid, "runs": reps * n,
                                          "launches": l0}
        sim.close()
    return out


def bench_rmsc01(args, rank, local_rank, dev, stream, sp, subscriptions=False):
    """config/rmsc01.py population (MarketMakerAgent, ZI, HBL over QUERY_ORDER_STREAM, Momentum); the first half hour of the day of every
    environment in one abx_run_kernel launch.  Reset untimed; one warm-up run, one timed run on fresh seeds."""
    import torch
    from marl_optimal_execution_b200 import _lib, distributed as D
    from marl_optimal_execution_b200.sim import BatchedSim, rmsc01_config, rmsc02_config
    n = args.rmsc01_envs_per_gpu
    sim = BatchedSim(rmsc02_config() if subscriptions else rmsc01_config(stop_ns=10 * 3600 * NS), n, device=local_rank)      # rmsc02: the whole day (~1.2e5 messages)
    res = None
    for rep in range(2):
        sim.reset(D.env_seeds(args.seed + 700001 * rep, rank * n, (rank + 1) * n), stream=sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier(); torch.cuda.synchronize(dev)
        lc = sim.launch_count
        e0.record(stream); sim.run(stream=sp); e1.record(stream)
        torch.cuda.synchronize(dev); D.barrier()
        if rep == 1:
            st = sim.stats(stream=sp)
            res = {"msgs": int(st["messages"].sum()), "ms": e0.elapsed_time(e1), "errs": int(((st["flags"] & _lib.F_ERROR_MASK) != 0).sum()),
                   "hbl_orders": int(st["limit_orders"].sum()), "launches": sim.launch_count - lc}
    sim.close()
    return res


def bench_marketreplay(args, rank, local_rank, dev, stream, sp):
    """BASELINE configs[4]: config/marketreplay.py -- the LOBSTER order stream of one day replayed through the books (no RL agent: order_level 0); one
    abx_env_step_kernel launch runs the whole day of every environment.  One untimed warm-up day, then a timed one."""
    import numpy as np
    import torch
    from marl_optimal_execution_b200 import _lib, distributed as D
    from marl_optimal_execution_b200.env import ABIDESEnv, env_config
    with np.load(os.path.join(ROOT, "tests", "golden", "mr_GOOG_2012-06-21.npz")) as g:
        stream5 = g["stream"].copy()
    n = args.mr_envs_per_gpu
    env = ABIDESEnv(stream5, n_envs=n, device=local_rank, cfg=env_config(order_level=0, stop_ns=(16 * 3600 + 60) * NS, queue_cap=256, level_cap=1024))
    env.reuse_outputs = True
    acts = torch.zeros(n, 3, dtype=torch.float64, device=dev)
    ms, msgs, errs, launches = 0.0, 0, 0, 0
    for rep in range(2):
        env.reset(stream=sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier(); torch.cuda.synchronize(dev)
        l0 = env.launch_count
        e0.record(stream); env.step(acts, stream=sp); e1.record(stream)
        torch.cuda.synchronize(dev); D.barrier()
        if rep == 1:
            st = env.stats(stream=sp)
            ms, msgs, launches = e0.elapsed_time(e1), int(st["messages"].sum()), env.launch_count - l0
            errs = int(((st["flags"] & _lib.F_ERROR_MASK) != 0).sum())
    env.close()
    return {"msgs": msgs, "ms": ms, "errs": errs, "launches": int(launches)}


def reference_rmsc03_block(cores):
    n_runs = max(8 * cores, 32)
    oracle_msgs_per_s(min(cores, 4), cores, 3)
    m, w, cpu_s = oracle_msgs_per_s(n_runs, cores, 3)
    return {"metric": "LOB msgs/sec (rmsc03)", "value": m / w, "unit": "msgs/s", "cores": cores, "kind": "port", "single_thread_value": m / cpu_s,
            "sample": "%d whole rmsc03 runs (09:30-09:46, no execution agent), event loop only, %d threads, %.1f CPU-s" % (n_runs, cores, cpu_s)}


def bench_env(args, rank, local_rank, dev, stream, sp):
    import numpy as np
    import torch
    from marl_optimal_execution_b200 import _lib, distributed as D
    from marl_optimal_execution_b200.env import ABIDESEnv

    n, K, W = args.env_envs_per_gpu, args.env_steps, max(args.warmup, 3)
    env = ABIDESEnv(replay_days(), n_envs=n, device=local_rank)
    env.reuse_outputs = True                              # no allocation inside the timed loops
    env.reset(stream=sp)
    gen = torch.Generator(device=dev); gen.manual_seed(args.seed + rank)
    acts = torch.rand(W + K + 1, n, 3, dtype=torch.float64, device=dev, generator=gen)
    acts[:, :, 0] *= 0.04
    for k in range(W + 1):                                # first step replays 00:00 -> 09:40 (13.5 k messages/env), untimed
        env.step(acts[k], stream=sp)
    torch.cuda.synchronize(dev)
    m0 = int(env.stats(stream=sp)["messages"].sum()); l0 = env.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier(); torch.cuda.synchronize(dev)
    e0.record(stream)
    for k in range(K):
        env.step(acts[W + 1 + k], stream=sp)
    e1.record(stream)
    torch.cuda.synchronize(dev); D.barrier()
    launches = env.launch_count - l0
    st = env.stats(stream=sp)
    m1 = int(st["messages"].sum())
    # e2e: host (pinned) buffers through abx_env_step_host: actions H2D, obs/reward/done D2H every step; a fresh episode (the timed
    # window above may have used most of the day), same untimed start-up
    env.reset(stream=sp)
    for k in range(W + 1):
        env.step(acts[k], stream=sp)
    a_pin = torch.rand(n, 3, dtype=torch.float64).pin_memory(); a_pin[:, 0] *= 0.04
    o_pin = torch.empty(n, 9, dtype=torch.float64).pin_memory(); r_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    d_pin = torch.empty(n, dtype=torch.uint8).pin_memory()
    D.barrier(); torch.cuda.synchronize(dev)
    w0 = time.perf_counter()
    for k in range(K):
        env.step_host_buffers(a_pin.data_ptr(), o_pin.data_ptr(), r_pin.data_ptr(), d_pin.data_ptr(), stream=sp)
    e2e_s = time.perf_counter() - w0
    D.barrier()
    errs = int(((env.stats(stream=sp)["flags"] & _lib.F_ERROR_MASK) != 0).sum())
    env.close()
    return {"steps": n * K, "msgs": m1 - m0, "ms": e0.elapsed_time(e1), "e2e_steps": n * K, "e2e_s": e2e_s, "errs": errs, "launches": int(launches)}


def bench_ddqn(args, rank, local_rank, dev, stream, sp):
    """DDQN acting loop: per decision tick one tcgen05 Q-network forward over all environments + one environment step."""
    import numpy as np
    import torch
    from marl_optimal_execution_b200 import _lib, distributed as D
    from marl_optimal_execution_b200.env import DDQNExecutionEnv
    from marl_optimal_execution_b200.qnet import QNetwork

    n, K, W = args.ddqn_envs_per_gpu, args.ddqn_steps, max(args.warmup, 3)
    env = DDQNExecutionEnv(replay_days(), n_envs=n, device=local_rank)
    env.reuse_outputs = True
    net = QNetwork(device=local_rank, seed=args.seed % 1000)
    qbuf = (None, torch.empty(n, dtype=torch.int32, device=dev))
    dq_seeds = np.arange(rank * n, (rank + 1) * n, dtype=np.uint64) + np.uint64(args.seed)

    def start():
        """fresh episodes brought to the first timed tick (untimed): 00:00 -> 10:00 start-up (~30 k messages/env) + W warm-up ticks; the device-timed and the
        e2e legs both start here, on the same seeds and the same action stream, so they cover the same simulated interval"""
        env.reset(seeds=dq_seeds, stream=sp)
        o, tr, rw, dn = env.step(torch.zeros(n, dtype=torch.int32, device=dev), stream=sp)
        tk = 0
        for _ in range(W):
            _, a_ = net.forward(o, x_offset=6, want_q=False, greedy_prob=0.9, seed=args.seed, counter=tk, out=qbuf, stream=sp)
            o, tr, rw, dn = env.step(a_, stream=sp); tk += 1
        torch.cuda.synchronize(dev)
        return o, tk
    obs, tick = start()
    m0 = int(env.stats(stream=sp)["messages"].sum()); l0 = env.launch_count + net.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    qe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    D.barrier(); torch.cuda.synchronize(dev)
    e0.record(stream)
    for k in range(K):
        qe[k][0].record(stream)
        _, a = net.forward(obs, x_offset=6, want_q=False, greedy_prob=0.9, seed=args.seed, counter=tick, out=qbuf, stream=sp)
        qe[k][1].record(stream)
        obs, trans, rew, done = env.step(a, stream=sp); tick += 1
    e1.record(stream)
    torch.cuda.synchronize(dev); D.barrier()
    launches = env.launch_count + net.launch_count - l0
    m1 = int(env.stats(stream=sp)["messages"].sum())
    qnet_ms = statistics.mean(a_.elapsed_time(b_) for a_, b_ in qe)
    # e2e: the same loop with every tick's results read back to pinned host memory (what a host-side learner consumes)
    o_pin = torch.empty(n, 8, dtype=torch.float64).pin_memory(); t_pin = torch.empty(n, 6, dtype=torch.float64).pin_memory()
    r_pin = torch.empty(n, dtype=torch.float64).pin_memory(); d_pin = torch.empty(n, dtype=torch.uint8).pin_memory()
    obs, tick = start()
    m2 = int(env.stats(stream=sp)["messages"].sum())
    D.barrier(); torch.cuda.synchronize(dev)
    w0 = time.perf_counter()
    for k in range(K):
        _, a = net.forward(obs, x_offset=6, want_q=False, greedy_prob=0.9, seed=args.seed, counter=tick, out=qbuf, stream=sp)
        obs, trans, rew, done = env.step(a, stream=sp); tick += 1
        o_pin.copy_(obs, non_blocking=True); t_pin.copy_(trans, non_blocking=True); r_pin.copy_(rew, non_blocking=True); d_pin.copy_(done, non_blocking=True)
        torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - w0
    D.barrier()
    assert int(env.stats(stream=sp)["messages"].sum()) - m2 == m1 - m0, "e2e and device-timed DDQN legs must cover the same messages"
    # training loop (marketreplay_ddqn_train shape): act, step, store the experience tuples, learn every 5 ticks on a batch from the shared
    # replay buffer (the reference's update rule, ddqn.py) and push the new weights into the acting network
    from marl_optimal_execution_b200.ddqn import DDQNTrainer
    tr = DDQNTrainer(device=dev, batch_size=args.ddqn_batch, seed=args.seed % 1000, buffer_capacity=max(4 * n * 8, 1 << 16))
    tr.agree_on_shard(n)
    net.set_params_device(tr.eval_net.flat_device())
    for _ in range(6):                                    # fill the buffer / warm the autograd kernels, untimed
        _, a = net.forward(obs, x_offset=6, want_q=False, greedy_prob=tr.greedy_prob(), seed=args.seed, counter=tick, out=qbuf, stream=sp)
        obs, trans, rew, done = env.step(a, stream=sp); tick += 1
        tr.store(trans)
    tr.learn()
    D.barrier(); torch.cuda.synchronize(dev)
    w0 = time.perf_counter(); l_before = tr.learn_step_counter
    for k in range(K):
        _, a = net.forward(obs, x_offset=6, want_q=False, greedy_prob=tr.greedy_prob(), seed=args.seed, counter=tick, out=qbuf, stream=sp)
        obs, trans, rew, done = env.step(a, stream=sp); tick += 1
        tr.store(trans)
        if k % tr.train_every == 0 and tr.learn() is not None:
            net.set_params_device(tr.eval_net.flat_device())
    torch.cuda.synchronize(dev)
    train_s = time.perf_counter() - w0
    D.barrier()
    st = env.stats(stream=sp)
    errs = int(((st["flags"] & _lib.F_ERROR_MASK) != 0).sum()) + int(d_pin.sum())       # no environment may have ended inside the device-timed / e2e ticks
    train_done = int(done.sum())       # the learner's policy may buy the parent order out before 15:30 (large scales): finished episodes, not errors
    env.close(); net.close()
    return {"steps": n * K, "msgs": m1 - m0, "ms": e0.elapsed_time(e1), "e2e_steps": n * K, "e2e_s": e2e_s, "errs": errs, "launches": int(launches), "qnet_ms": qnet_ms,
            "train_s": train_s, "learn_steps": tr.learn_step_counter - l_before, "buffer": tr.buffer.size, "train_done": train_done}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=16384)
    ap.add_argument("--slice-s", type=float, default=300.0)
    ap.add_argument("--variant", type=int, default=1000, choices=[100, 1000])
    ap.add_argument("--seed", type=int, default=123456789)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-env", action="store_true", help="skip the ABIDESEnv steps/s measurement")
    ap.add_argument("--env-envs-per-gpu", type=int, default=9472, help="4 x 148 SMs x 16 resident one-warp CTAs: whole waves (8192 leaves the 4th wave 46 %% full)")
    ap.add_argument("--env-steps", type=int, default=750, help="timed ABIDESEnv steps: 750 = the whole 761-tick episode after the start-up and warm-up steps")
    ap.add_argument("--no-rmsc03", action="store_true", help="skip the rmsc03 population (BASELINE configs[2])")
    ap.add_argument("--no-rmsc01", action="store_true", help="skip the rmsc01 population (MarketMaker + ZI + HBL + Momentum)")
    ap.add_argument("--rmsc01-envs-per-gpu", type=int, default=2368, help="148 SMs x 16 resident one-warp CTAs")
    ap.add_argument("--no-whole-day", action="store_true", help="skip the whole-environment-day measurement of the headline workload")
    ap.add_argument("--no-marketreplay", action="store_true", help="skip the config/marketreplay.py shape (BASELINE configs[4])")
    ap.add_argument("--mr-envs-per-gpu", type=int, default=4736, help="2 x 148 SMs x 16 resident one-warp CTAs")
    ap.add_argument("--rmsc03-envs-per-gpu", type=int, default=4096, help="BASELINE configs[2]: 4096 envs/GPU")
    ap.add_argument("--no-ddqn", action="store_true", help="skip the DDQN execution shape (Q-network forward + environment step per tick)")
    ap.add_argument("--ddqn-envs-per-gpu", type=int, default=9472, help="whole waves of 148 x 16 resident environments, like --env-envs-per-gpu")
    ap.add_argument("--ddqn-steps", type=int, default=200, help="timed DDQN decision ticks per sub-measurement (acting, e2e, training share one 660-tick day)")
    ap.add_argument("--ddqn-batch", type=int, default=4096, help="learner batch size of the DDQN training sub-measurement")
    args = ap.parse_args()
    quiet_stdout()
    if args.warmup < 3 and args.impl == "ours":
        print("bench.py: note: W >= 3 is required for a valid number", file=sys.stderr)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    from marl_optimal_execution_b200 import distributed as D
    rank, local_rank, world = D.init()
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
