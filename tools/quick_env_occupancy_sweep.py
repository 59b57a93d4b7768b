"""ABIDESEnv steps/s against the number of environments (whole IBM day, one stream): how throughput follows resident warps per SM."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from marl_optimal_execution_b200.env import ABIDESEnv
dev = torch.device('cuda', 0)
g = np.load('/root/repo/tests/golden/env_IBM_2003-01-14_s789.npz')
for n in [int(a) for a in sys.argv[1:]] or [592, 1184, 2368, 3404, 4736, 8192]:
    env = ABIDESEnv(g['stream'], n_envs=n); env.reset()
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    acts = [torch.rand(n, 3, dtype=torch.float64, device=dev, generator=gen) * torch.tensor([0.04, 1, 1], dtype=torch.float64, device=dev) for _ in range(8)]
    for k in range(40): env.step(acts[k % 8])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(300): env.step(acts[k % 8])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("n_envs %6d (%.1f warps/SM): %.3f ms/step, %.4g steps/s" % (n, n / 148, 1e3 * dt / 300, 300 * n / dt), flush=True)
    del env
