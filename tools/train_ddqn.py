#!/usr/bin/env python3
"""The reference's DDQN training shape on the batched simulator (config/execution/marketreplay/execution_marketreplay_ddqn_parallel.py:40-75,
marketreplay_ddqn_train): the nine train dates 2003-01-13..17 and 2003-01-21..24, BUY 5e5 from 10:00 over 330 minutes in 30 s ticks, ONE policy carried
from episode to episode.  Environment e starts on day e % 9 and moves on to its next day after every episode (auto-reset), so every episode of the batch
covers all nine dates; with several GPUs (torchrun) the ranks train one policy (gradient all-reduce over NCCL).

    python tools/train_ddqn.py [--envs-per-gpu 2368] [--episodes 9] [--out profiles/r02_ddqn_learning_curve.json]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_ddqn.py ...

Writes one JSON document: the learning curve (per episode: mean / std total step reward over all environments, mean loss, learn steps, ticks/s)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_optimal_execution_b200 import distributed as D                        # noqa: E402
from marl_optimal_execution_b200.ddqn import DDQNTrainer                        # noqa: E402
from marl_optimal_execution_b200.env import DDQNExecutionEnv                    # noqa: E402
from marl_optimal_execution_b200.qnet import QNetwork                           # noqa: E402

TRAIN_DATES = ["2003-01-13", "2003-01-14", "2003-01-15", "2003-01-16", "2003-01-17", "2003-01-21", "2003-01-22", "2003-01-23", "2003-01-24"]
GOLDEN = {"2003-01-14": "env_IBM_2003-01-14_s789.npz", "2003-01-15": "env_IBM_2003-01-15_s4242.npz", "2003-01-16": "ddqn_IBM_2003-01-16_s99_sell.npz"}


def train_days():
    out = []
    for d in TRAIN_DATES:
        f = os.path.join(ROOT, "tests", "golden", GOLDEN[d]) if d in GOLDEN else os.path.join(ROOT, "tests", "golden", "days", "IBM_%s.npz" % d)
        with np.load(f) as g:
            out.append(g["stream"].copy())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=2368)
    ap.add_argument("--episodes", type=int, default=9)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--lr", type=float, default=0.001)
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_ddqn_learning_curve.json"))
    a = ap.parse_args()
    rank, local_rank, world = D.init()
    dev = torch.device("cuda", local_rank)
    n = a.envs_per_gpu
    env = DDQNExecutionEnv(train_days(), n_envs=n, device=local_rank)
    env.reuse_outputs = True
    env.reset(seeds=np.arange(rank * n, (rank + 1) * n, dtype=np.uint64) + np.uint64(a.seed))
    net = QNetwork(device=local_rank, seed=a.seed)
    tr = DDQNTrainer(device=dev, batch_size=a.batch, learning_rate=a.lr, seed=a.seed, epsilon_increment=0.9 / 400, buffer_capacity=1 << 21)
    net.set_params_device(tr.eval_net.flat_device())
    qbuf = (None, torch.empty(n, dtype=torch.int32, device=dev))
    state = {"tick": 0}

    def act(obs, greedy_prob, tick):
        state["tick"] += 1
        _, act_ = net.forward(obs, x_offset=6, want_q=False, greedy_prob=greedy_prob, seed=a.seed + rank, counter=state["tick"], out=qbuf)
        return act_

    def sync(_flat):
        net.set_params_device(tr.eval_net.flat_device())

    tr.eval_net.flat = tr.eval_net.flat_device                                 # run_episode hands sync_fn the flat parameters: keep them on the device
    curve, t0 = [], time.perf_counter()

    def on_episode(rec):
        torch.cuda.synchronize(dev)
        g = D.gather_summaries(torch.tensor([rec["mean_total_reward"] * n, n, rec["ticks"] * n], dtype=torch.float64), device=dev)
        rec = dict(rec, mean_total_reward=float(g[:, 0].sum() / g[:, 1].sum()), envs=int(g[:, 1].sum()), wall_s=time.perf_counter() - t0,
                   env_ticks_per_s=float(g[:, 2].sum()) / max(time.perf_counter() - on_episode.t_last, 1e-9))
        on_episode.t_last = time.perf_counter()
        curve.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
    on_episode.t_last = time.perf_counter()
    tr.run_episodes(env, act, a.episodes, sync_fn=sync, on_episode=on_episode)
    if rank == 0:
        doc = {"what": "DDQN training sweep over the reference's nine train dates (IBM LOBSTER sample days), BUY 5e5 10:00 + 330 min, 30 s ticks; one policy, "
                       "%d environments on %d GPU(s); total step reward = sum of compute_reward over an episode (ddqlearning_execution_agent.py:411-447)" % (n * world, world),
               "args": vars(a), "n_gpus": world, "curve": curve}
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(doc, open(a.out, "w"), indent=1)
    env.close(); net.close()
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
