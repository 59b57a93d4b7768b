"""Quick GPU sanity + timing probe (development aid, run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.sim import BatchedSim, sparse_zi_config

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cfg = sparse_zi_config(variant)
sim = BatchedSim(cfg, n_envs)
print("device bytes %.2f GB" % (sim.device_bytes / 1e9))
sim.reset(np.arange(n_envs, dtype=np.uint64) + 1)
torch.cuda.synchronize()
NS = 10 ** 9
t_open = int(cfg.mkt_open_ns)
marks = [t_open - NS, t_open + 600 * NS, t_open + 1800 * NS, t_open + 3600 * NS]
prev = 0
for m in marks:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.run(m); e1.record(); torch.cuda.synchronize()
    st = sim.stats()
    tot = int(st["messages"].sum())
    ms = e0.elapsed_time(e1)
    print("until %6.0fs: %.1f ms, +%d msgs, %.3e msgs/s, flags %s maxq %d lv %d/%d rest %d" % (
        m / NS, ms, tot - prev, (tot - prev) / (ms / 1e3), hex(int(np.bitwise_or.reduce(st["flags"]))), st["max_queue"].max(),
        st["n_bid_levels"].max(), st["n_ask_levels"].max(), st["n_resting"].max()))
    prev = tot
