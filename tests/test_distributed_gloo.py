"""N>1 host logic on CPU: world_size-2 gloo processes shard the environment index space and all-gather episode
statistics (the only collective on this path)."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from marl_optimal_execution_b200 import _lib, distributed as D
    r, _, w = D.init("gloo")
    lo, hi = D.shard_range(11, r, w)
    stats = np.zeros(hi - lo, dtype=_lib.STATS_DTYPE)
    stats["messages"] = np.arange(lo, hi) + 100
    stats["fills"] = 1
    stats["flags"][:1] = _lib.F_QUEUE_OVERFLOW if r == 1 else 0
    D.barrier()
    g = D.gather_summaries(D.summarize(stats))
    mx = D.max_over_ranks(10.0 + r)
    q.put((r, lo, hi, g.tolist(), mx, D.env_seeds(1000, lo, hi).tolist()))
    torch.distributed.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    (r0, lo0, hi0, g0, mx0, s0), (r1, lo1, hi1, g1, mx1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 6, 6, 11)                 # contiguous, exhaustive, balanced
    assert s0 + s1 == list(range(1000, 1011))                    # seed = base + global env index
    assert g0 == g1 and mx0 == mx1 == 11.0
    assert g0[0][0] == sum(range(100, 106)) and g0[1][0] == sum(range(106, 111))
    assert [row[5] for row in g0] == [0, 1] and [row[6] for row in g0] == [6, 5] and [row[3] for row in g0] == [6, 5]


def test_shard_range_properties():
    from marl_optimal_execution_b200.distributed import shard_range
    for n in (1, 7, 8, 16384, 65536):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _learner(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from marl_optimal_execution_b200 import distributed as D
    from marl_optimal_execution_b200.ddqn import DDQNTrainer
    D.init("gloo")
    tr = DDQNTrainer(batch_size=8, seed=5, buffer_capacity=256)             # same seed: identical initial weights on every rank
    rs = np.random.RandomState(100 + rank)                                  # different experience per rank
    t = np.column_stack([rs.randint(0, 200, (64, 2)), rs.randint(0, 24, 64), rs.randint(0, 200, (64, 2)), rs.uniform(0, 20, 64)])
    tr.buffer.push(torch.from_numpy(t))
    for _ in range(3):
        tr.learn()
    q.put((rank, tr.eval_net.flat().tolist(), float(tr.cost_hist[0])))
    torch.distributed.destroy_process_group()


def test_one_policy_across_ranks_gradient_allreduce_world2():
    """The only collective of the training path: ranks hold different experience, average their gradients, and end with identical weights."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    ps = [ctx.Process(target=_learner, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=180) for _ in ps)
    [p.join(60) for p in ps]
    (_, w0, c0), (_, w1, c1) = res
    assert w0 == w1 and c0 != c1                                            # same policy, different local batches
