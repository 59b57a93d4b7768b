"""ABIDESEnv / DDQN step time along the trading day (ms per step for all environments, every 40 steps)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from marl_optimal_execution_b200.env import ABIDESEnv, DDQNExecutionEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device('cuda', 0)
g = np.load('/root/repo/tests/golden/env_IBM_2003-01-14_s789.npz')
env = ABIDESEnv(g['stream'], n_envs=n); env.reset()
gen = torch.Generator(device=dev); gen.manual_seed(1)
rows = []
k = 0; done = torch.zeros(n, dtype=torch.uint8, device=dev)
m_prev = 0
while k < 761:
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(40):
        a = torch.rand(n, 3, dtype=torch.float64, device=dev, generator=gen); a[:, 0] *= 0.04
        obs, rew, done, _ = env.step(a); k += 1
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = env.stats(); m = int(st["messages"].sum())
    rows.append((k, 1e3 * dt / 40, (m - m_prev) / n / 40, int(st["n_resting"].max()), int(st["n_bid_levels"].max()), int(done.sum())))
    m_prev = m
print("ABIDESEnv n=%d: (step, ms/step, msgs/env/step, max resting, max bid levels, done)" % n)
for r in rows: print("  %4d %8.3f %8.1f %6d %5d %5d" % r)
tot = sum(r[1] * 40 for r in rows) / 1e3
print("episode: %.2f s for %d steps x %d envs = %.3g steps/s" % (tot, k, n, k * n / tot))
