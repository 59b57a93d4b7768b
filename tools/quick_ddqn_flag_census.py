"""Which environments does bench.py's DDQN section count in `error_envs`, and when?  Runs the bench's acting loop (3 IBM days round robin,
epsilon-greedy 0.9 actions from the random-init network) for the same 609 ticks and prints, every 100 ticks, the histogram of status flags,
the number of finished environments and which replayed day they belong to."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.env import DDQNExecutionEnv
from marl_optimal_execution_b200.qnet import QNetwork

n = 9472; dev = torch.device("cuda", 0); seed = 123456789
days = bench.replay_days()
env = DDQNExecutionEnv(days, n_envs=n, device=0)
net = QNetwork(device=0, seed=seed % 1000)
env.reset(seeds=np.arange(n, dtype=np.uint64) + np.uint64(seed))
obs, trans, rew, done = env.step(torch.zeros(n, dtype=torch.int32, device=dev))
ever_done = torch.zeros(n, dtype=torch.bool, device=dev)
first_done = torch.full((n,), -1, dtype=torch.int32, device=dev)
for tick in range(609):
    _, a = net.forward(obs, x_offset=6, want_q=False, greedy_prob=0.9, seed=seed, counter=tick)
    obs, trans, rew, done = env.step(a)
    d = done.bool()
    first_done = torch.where(d & ~ever_done, torch.full_like(first_done, tick), first_done)
    ever_done |= d
    if tick % 100 == 99 or tick == 608:
        fl = env.stats()["flags"]
        err = (fl & _lib.F_ERROR_MASK) != 0
        vals, cnt = np.unique(fl, return_counts=True)
        print("tick", tick + 1, "flags", {hex(int(v)): int(c) for v, c in zip(vals, cnt)}, "error envs", int(err.sum()),
              "done now", int(d.sum()), "ever done", int(ever_done.sum()), flush=True)
fl = env.stats()["flags"]
bad = np.nonzero(((fl & _lib.F_ERROR_MASK) != 0) | ever_done.cpu().numpy())[0]
fd = first_done.cpu().numpy()
print("environments counted:", len(bad))
for e in bad[:40]:
    print("  env", int(e), "day", int(e) % len(days), "flags", hex(int(fl[e])), "first done at tick", int(fd[e]))
