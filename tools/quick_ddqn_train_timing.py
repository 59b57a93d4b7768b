import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from marl_optimal_execution_b200.env import DDQNExecutionEnv
from marl_optimal_execution_b200.qnet import QNetwork
from marl_optimal_execution_b200.ddqn import DDQNTrainer
g = np.load('/root/repo/tests/golden/ddqn_IBM_2003-01-14_s4242.npz')
n = 9472; dev = torch.device('cuda', 0)
env = DDQNExecutionEnv(g['stream'], n_envs=n); env.reset(seeds=np.arange(n, dtype=np.uint64))
net = QNetwork(seed=1)
tr = DDQNTrainer(device=dev, batch_size=4096, seed=1, buffer_capacity=1 << 18)
obs, trans, rew, done = env.step(torch.zeros(n, dtype=torch.int32, device=dev))
def sync(): torch.cuda.synchronize()
T = {k: 0.0 for k in ('fwd', 'step', 'push', 'learn', 'setp')}
for k in range(46):
    sync(); t0 = time.perf_counter(); _, a = net.forward(obs, x_offset=6, want_q=False, greedy_prob=0.5, seed=1, counter=k); sync(); t1 = time.perf_counter()
    obs, trans, rew, done = env.step(a); sync(); t2 = time.perf_counter()
    tr.buffer.push(trans); sync(); t3 = time.perf_counter()
    if k % 5 == 0:
        tr.learn(); sync(); t4 = time.perf_counter(); net.set_params_device(tr.eval_net.flat_device()); sync(); t5 = time.perf_counter()
    else: t4 = t5 = t3
    if k >= 6:
        T['fwd'] += t1 - t0; T['step'] += t2 - t1; T['push'] += t3 - t2; T['learn'] += t4 - t3; T['setp'] += t5 - t4
print({k: round(1e3 * v / 40, 3) for k, v in T.items()}, "ms per tick (40 ticks)")
# the same loop without per-component synchronisation (as bench.py times it)
sync(); w0 = time.perf_counter()
for k in range(40):
    _, a = net.forward(obs, x_offset=6, want_q=False, greedy_prob=tr.greedy_prob(), seed=1, counter=100 + k)
    obs, trans, rew, done = env.step(a)
    tr.buffer.push(trans)
    if k % 5 == 0 and tr.learn() is not None:
        net.set_params_device(tr.eval_net.flat_device())
sync(); print("unsynced loop: %.3f ms per tick" % (1e3 * (time.perf_counter() - w0) / 40))
sync(); w0 = time.perf_counter()
for k in range(40):
    _, a = net.forward(obs, x_offset=6, want_q=False, greedy_prob=0.9, seed=1, counter=200 + k)
    obs, trans, rew, done = env.step(a)
sync(); print("acting only: %.3f ms per tick" % (1e3 * (time.perf_counter() - w0) / 40), "flags", np.unique(env.stats()["flags"]), "done", int(done.sum()))
