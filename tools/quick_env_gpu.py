"""Quick GPU timing probe of the ABIDESEnv path (development aid, run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from marl_optimal_execution_b200.env import ABIDESEnv, env_config
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "env_IBM_2003-01-14_s789.npz"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
env = ABIDESEnv(g["stream"], n_envs=n)
env.reset()
rng = np.random.RandomState(1)
acts = torch.from_numpy(np.stack([rng.uniform(0, 0.04, (steps + 1, n)), rng.uniform(0, 1, (steps + 1, n)), rng.uniform(0, 1, (steps + 1, n))], -1)).cuda()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); env.step(acts[0]); e1.record(); torch.cuda.synchronize()
m0 = int(env.stats()["messages"].sum())
print("first step (00:00 -> 09:40): %.1f ms, %d msgs, %.3e msgs/s" % (e0.elapsed_time(e1), m0, m0 / (e0.elapsed_time(e1) / 1e3)))
e0.record()
for k in range(1, steps + 1):
    obs, rew, done, _ = env.step(acts[k])
e1.record(); torch.cuda.synchronize()
st = env.stats(); m1 = int(st["messages"].sum())
ms = e0.elapsed_time(e1)
print("%d envs x %d steps: %.1f ms -> %.3e steps/s, %.3e msgs/s, flags %s, msgs/step %.1f" % (n, steps, ms, n * steps / (ms / 1e3), (m1 - m0) / (ms / 1e3), hex(int(np.bitwise_or.reduce(st["flags"]))), (m1 - m0) / (n * steps)))
