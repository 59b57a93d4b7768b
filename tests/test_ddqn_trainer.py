"""Host logic of the DDQN learner (marl_optimal_execution_b200/ddqn.py) on the CPU: the update must equal a numpy restatement of
train_neural_nets (agent/execution/qlearning/ddqlearning_execution_agent.py:449-505) with Keras' RMSprop rule, and the acting /
storing / learning loop must run over the batched environment (host emulation harness of the simulator logic; a test tool)."""
import os

import numpy as np
import torch

from helpers import build_emu
from marl_optimal_execution_b200 import _lib
from marl_optimal_execution_b200.ddqn import DDQNTrainer, ReplayBuffer, TorchMLP
from marl_optimal_execution_b200.env import DDQNExecutionEnv, dq_config
from marl_optimal_execution_b200.qnet import DEFAULT_DIMS, init_params, unpack_params


def np_forward(flat, x, dims=DEFAULT_DIMS):
    h = x.astype(np.float32)
    layers = unpack_params(flat, dims)
    for i, (w, b) in enumerate(layers):
        h = h @ w.T + b
        if i + 1 < len(layers):
            h = np.maximum(h, 0)
    return h


def test_learn_step_equals_reference_update_rule():
    tr = DDQNTrainer(batch_size=16, seed=3, buffer_capacity=64)
    tr.eval_net.DROPOUT = 0.0                                             # the exact update rule first; the training-time Dropout(0.1) has its own test below
    rs = np.random.RandomState(0)
    t = np.column_stack([rs.randint(0, 200, (40, 2)), rs.randint(0, 24, 40), rs.randint(0, 200, (40, 2)), rs.uniform(0, 20, 40)])
    t[5, 5] = np.nan                                                      # r None: skipped
    assert tr.buffer.push(torch.from_numpy(t)) == 40 and tr.buffer.n_valid() == 39
    # replicate the sampling, then the reference arithmetic in numpy
    g = torch.Generator(); g.manual_seed(3)
    valid = torch.ones(40); valid[5] = 0
    idx = torch.multinomial(valid + 1e-30, 16, replacement=True, generator=g).numpy()
    assert 5 not in idx
    ok = t[idx]
    s, a, sp, r = ok[:, 0:2], ok[:, 2].astype(int), ok[:, 3:5], ok[:, 5].astype(np.float32)
    ev0, tg0 = tr.eval_net.flat(), tr.target_net.flat()
    q_next = np_forward(tg0, sp)
    tgt_a = r + np.float32(0.98) * q_next[np.arange(16), q_next.argmax(1)]     # r + gamma * q_next[argmax q_eval4next], both from the TARGET net
    ref = TorchMLP(DEFAULT_DIMS, flat=ev0)
    s_t = torch.from_numpy(s.astype(np.float32))
    q_target = ref(s_t).detach().clone()                                      # q_eval.copy(): RMSprop normalises gradients, so the untouched
    assert np.allclose(q_target.numpy(), np_forward(ev0, s), rtol=1e-4, atol=1e-3)   # entries must cancel exactly, as in train_on_batch
    new_a = torch.from_numpy(tgt_a)
    assert np.allclose(tr.target_net(torch.from_numpy(sp.astype(np.float32))).detach().numpy(), q_next, rtol=1e-4, atol=1e-3)
    cost = tr.learn()
    # learn_step_counter 0 -> the target network was replaced by the eval network BEFORE the fit (:487-490)
    assert np.array_equal(tr.target_net.flat(), ev0)
    # MSE gradient + Keras RMSprop: v = 0.9 v + 0.1 g^2 ; w -= lr g / (sqrt(v) + 1e-7)
    tn = TorchMLP(DEFAULT_DIMS, flat=tg0)
    with torch.no_grad():
        qn = tn(torch.from_numpy(sp.astype(np.float32)))
        q_target[torch.arange(16), torch.from_numpy(a)] = torch.from_numpy(r) + 0.98 * qn.gather(1, qn.argmax(1)[:, None])[:, 0]
    assert np.allclose(q_target[torch.arange(16), torch.from_numpy(a)].numpy(), new_a.numpy(), rtol=1e-4, atol=1e-3)
    loss = ((ref(s_t) - q_target) ** 2).mean()
    loss.backward()
    cost = float(cost)
    assert abs(float(loss.detach()) - cost) <= 1e-5 * max(1.0, abs(cost))
    for p, q in zip(ref.params, tr.eval_net.params):
        gnp = p.grad.numpy()
        want = p.detach().numpy() - 0.01 * gnp / (np.sqrt(0.1 * gnp * gnp) + 1e-7)
        assert np.allclose(q.detach().numpy(), want, rtol=1e-4, atol=1e-6)
    assert tr.learn_step_counter == 1 and tr.epsilon == 0.9


def test_training_forward_applies_keras_dropout():
    """util/model/QNets.py:16-25: Dropout(0.1) after hidden layers 2..6 in train_on_batch (inverted dropout: kept units / 0.9), identity in predict()."""
    net = TorchMLP(DEFAULT_DIMS, seed=5)
    x = torch.rand(4096, 2) * 200
    g = torch.Generator(); g.manual_seed(1)
    with torch.no_grad():
        plain, again, train = net(x), net(x, training=False), net(x, training=True, generator=g)
        assert torch.equal(plain, again) and not torch.equal(plain, train)
        # first hidden layer untouched, second one: ~10 % of the positive units zeroed, the others scaled by 1 / 0.9
        w1, b1, w2, b2 = (p.detach() for p in net.params[:4])
        h1 = torch.relu(x @ w1.T + b1)
        h2 = torch.relu(h1 @ w2.T + b2)
        g2 = torch.Generator(); g2.manual_seed(1)
        keep = (torch.rand(h2.shape, generator=g2) >= 0.1).float()
        assert abs(float(keep.mean()) - 0.9) < 0.01
        h2d = h2 * keep / 0.9
        assert abs(float(h2d.mean()) / float(h2.mean()) - 1.0) < 0.02          # expectation preserved
    # the learner's loss uses the training forward, its targets the inference forward
    tr = DDQNTrainer(batch_size=32, seed=2, buffer_capacity=256)
    tr.store(torch.rand(200, 6, dtype=torch.float64) * 20)
    assert tr.can_learn() and tr.learn() is not None and np.isfinite(float(tr.cost_hist[-1]))


def test_learner_with_nnmodel_2():
    """util/model/QNets.py:30-52 NNModel_2 (eight Dense layers, 256 wide in the middle, Dropout after hidden layers 2..7): the learner is generic over the layer list."""
    from marl_optimal_execution_b200.qnet import NNMODEL_2_DIMS, param_count
    assert param_count(NNMODEL_2_DIMS) == 87576
    tr = DDQNTrainer(dims=NNMODEL_2_DIMS, batch_size=32, seed=3, buffer_capacity=256)
    tr.store(torch.rand(200, 6, dtype=torch.float64) * 20)
    before = tr.eval_net.flat().copy()
    assert tr.can_learn() and tr.learn() is not None and np.isfinite(float(tr.cost_hist[-1]))
    assert tr.eval_net.flat().size == 87576 and not np.array_equal(before, tr.eval_net.flat())


def test_buffer_without_valid_rows_cannot_break_sampling():
    tr = DDQNTrainer(batch_size=4, seed=2, buffer_capacity=64)
    t = torch.rand(20, 6, dtype=torch.float64); t[:, 5] = float("nan")        # only r == None rows
    tr.store(t)
    before = tr.eval_net.flat()
    assert tr.learn() is not None and np.array_equal(tr.eval_net.flat(), before)   # zero-weight batch: no update, no exception


def test_replay_buffer_wraps_and_epsilon_schedule():
    b = ReplayBuffer(8, "cpu")
    t = torch.arange(60, dtype=torch.float64).reshape(10, 6)
    assert b.push(t) == 8 and b.size == 8
    assert set(b.r.tolist()) == {float(6 * i + 5) for i in range(2, 10)}
    tr = DDQNTrainer(batch_size=4, epsilon_increment=0.3, buffer_capacity=32)
    assert tr.epsilon == 0.0 and tr.greedy_prob() == 0.0
    tr.buffer.push(torch.rand(10, 6, dtype=torch.float64) * 20)
    for _ in range(5):
        tr.learn()
    assert abs(tr.epsilon - 0.9) < 1e-9 or tr.epsilon >= 0.9              # 0 -> .3 -> .6 -> .9 -> 1.2? no: clamps to epsilon_max once >= max
    assert tr.epsilon <= 1.2 + 1e-9


def test_training_loop_over_the_batched_environment(golden_dir):
    emu = build_emu()
    L = _lib.load(emu)
    g = np.load(os.path.join(golden_dir, "ddqn_IBM_2003-01-14_s4242.npz"))
    env = DDQNExecutionEnv(g["stream"], n_envs=4, cfg=dq_config(L), lib_path=emu)
    env.reset(seeds=np.arange(4, dtype=np.uint64))
    tr = DDQNTrainer(batch_size=8, seed=1, buffer_capacity=4096)
    pushed = []

    def act(obs, greedy_prob, tick):
        q = tr.eval_net(torch.as_tensor(obs[:, 6:8], dtype=torch.float32))
        a = q.argmax(dim=1).numpy().astype(np.int32)
        rnd = np.random.RandomState(tick).randint(0, 24, len(a)).astype(np.int32)
        return np.where(np.random.RandomState(1000 + tick).uniform(size=len(a)) < greedy_prob, a, rnd)

    total, ticks = tr.run_episode(env, act, sync_fn=lambda flat: pushed.append(flat), max_ticks=60)
    assert ticks == 60 and tr.buffer.size > 150 and tr.learn_step_counter >= 8 and len(pushed) == tr.learn_step_counter
    assert np.isfinite([float(c) for c in tr.cost_hist]).all() and float(total.min()) > 0.0           # every environment filled something: rewards are positive
    st = env.stats()
    assert (st["flags"] & _lib.F_ERROR_MASK == 0).all()
