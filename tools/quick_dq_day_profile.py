"""DDQN execution shape: ms per decision tick along the day (random actions)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from marl_optimal_execution_b200.env import DDQNExecutionEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device('cuda', 0)
g = np.load('/root/repo/tests/golden/ddqn_IBM_2003-01-14_s4242.npz')
env = DDQNExecutionEnv(g['stream'], n_envs=n); env.reset(seeds=np.arange(n, dtype=np.uint64))
gen = torch.Generator(device=dev); gen.manual_seed(1)
obs, trans, rew, done = env.step(torch.zeros(n, dtype=torch.int32, device=dev))
k, tot, rows = 0, 0.0, []
while k < 660:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(60):
        a = torch.randint(0, 24, (n,), dtype=torch.int32, device=dev, generator=gen)
        obs, trans, rew, done = env.step(a); k += 1
    torch.cuda.synchronize(); dt = time.perf_counter() - t0; tot += dt
    rows.append((k, 1e3 * dt / 60))
st = env.stats()
print("DDQN n=%d ms/tick per 60 ticks:" % n, ["%.2f" % r[1] for r in rows], "flags", np.unique(st["flags"]), "max_queue", int(st["max_queue"].max()))
print("day: %.2f s for 660 ticks x %d envs = %.3g ticks/s" % (tot, n, 660 * n / tot))
