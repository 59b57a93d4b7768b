#!/usr/bin/env python3
"""Small bit-exact parity runs for compute-sanitizer (racecheck / synccheck), no torch:

    compute-sanitizer --tool racecheck python tools/sanitize_parity.py [lib.so]
    compute-sanitizer --tool synccheck python tools/sanitize_parity.py [lib.so]

sparse_zi_100 (tape mode, 4 environments, whole day), rmsc03 (tape mode, first two simulated minutes), ABIDESEnv (30 steps) and the DDQN
shape (5 ticks), each checked against the oracle, so the sanitizer sees every on-chip structure (event queue tiers, ladders, staged records,
order cache) being exercised by a run whose results are known to be right."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import oracle_tapes                                              # noqa: E402
from marl_optimal_execution_b200 import _lib                                  # noqa: E402
from marl_optimal_execution_b200.env import ABIDESEnv, DDQNExecutionEnv, dq_config, env_config   # noqa: E402
from marl_optimal_execution_b200.sim import BatchedSim, rmsc03_config, sparse_zi_config          # noqa: E402
from oracle.oracle import OracleDDQNEnv, OracleEnv, OracleSim, TRACE_ALL      # noqa: E402

lib = sys.argv[1] if len(sys.argv) > 1 else None
NS = 10 ** 9

seeds = [123456789, 1001, 7, 8]
orc = [OracleSim(100, s, TRACE_ALL) for s in seeds]
cnt = [o.run() for o in orc]
sim = BatchedSim(sparse_zi_config(100, rng_mode=_lib.RNG_TAPE, hash_pops=1), 4, lib_path=lib)
sim.reset_tape(*oracle_tapes(orc))
sim.run(); sim.finalize()
st = sim.stats()
assert [int(x) for x in st["pop_hash"]] == [o.pop_hash() for o in orc] and list(st["messages"]) == cnt
sim.close()
print("sparse_zi_100 x4: ok", cnt)

o3 = OracleSim(3, 1001, TRACE_ALL); o3.run()
o3b = OracleSim(3, 1001, 0); n3, _ = o3b.run_until((9 * 3600 + 32 * 60) * NS)
sim = BatchedSim(rmsc03_config(rng_mode=_lib.RNG_TAPE, hash_pops=1), 2, lib_path=lib)
sim.reset_tape(*oracle_tapes([o3, o3]))
sim.run((9 * 3600 + 32 * 60) * NS)
st = sim.stats()
assert (st["pop_hash"] == np.uint64(o3b.pop_hash())).all() and (st["messages"] == n3).all()
sim.close()
print("rmsc03 x2 (two minutes): ok", n3)

g = np.load(os.path.join(ROOT, "tests", "golden", "env_IBM_2003-01-14_s789.npz"))
env = ABIDESEnv(g["stream"], n_envs=3, cfg=env_config(hash_pops=1), lib_path=lib)
env.reset(); oe = OracleEnv(g["stream"])
for k in range(30):
    obs, _, done, _ = env.step(np.tile(g["actions"][k], (3, 1)))
    oo, _, od, _ = oe.step(g["actions"][k])
    assert np.allclose(obs[0][: len(oo)], oo, rtol=1e-9, atol=1e-12)
assert (env.stats()["pop_hash"] == np.uint64(oe.pop_hash())).all()
env.close()
print("ABIDESEnv x3 (30 steps): ok", oe.n_pops)

g = np.load(os.path.join(ROOT, "tests", "golden", "ddqn_IBM_2003-01-14_s4242.npz"))
denv = DDQNExecutionEnv(g["stream"], n_envs=2, cfg=dq_config(hash_pops=1), lib_path=lib)
denv.reset(mom_sizes=np.tile(g["mom_sizes"].astype(np.int32), (2, 1)))
od_ = OracleDDQNEnv(g["stream"], g["mom_sizes"])
denv.step(None); od_.step(0)
for k in range(5):
    a = int(g["actions"][k]); denv.step(np.full(2, a, np.int32)); od_.step(a)
assert (denv.stats()["pop_hash"] == np.uint64(od_.pop_hash())).all()
denv.close()
print("DDQN shape x2 (5 ticks): ok", od_.n_pops)
