"""rmsc03 (+ POV execution agent) batch throughput: 4 096 environments per GPU, the whole 09:30-09:46 run in one launch (BASELINE configs[2])."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from marl_optimal_execution_b200.sim import BatchedSim, rmsc03_config
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for pov in (False, True):
    sim = BatchedSim(rmsc03_config(pov_exec=pov), n)
    for rep in range(2):
        sim.reset(np.arange(n, dtype=np.uint64) + 1000 * rep + 7)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sim.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = sim.stats()
    print("rmsc03%s: %d envs, %.0f messages/env, %.3f s -> %.4g msgs/s, flags %s" % (" + POV execution agent" if pov else "", n, st["messages"].mean(), dt, st["messages"].sum() / dt, np.unique(st["flags"])))
