"""Numerics of the tcgen05 Q-network forward (csrc/abx_qnet.cu) against a plain PyTorch fp32 reference of the same network
(util/model/QNets.py:7-27; TensorFlow is absent from the image, so the Keras arithmetic itself is unpinned -- SURVEY section 8c).
Tolerance: the kernel multiplies bf16 hi/lo splits on the tensor cores (three MMAs per k step, fp32 accumulation), which keeps
~16 mantissa bits per operand: |q - q_ref| <= 2e-4 * max|q_ref| + 1e-5; actions must equal np.argmax except at near-ties."""
import numpy as np
import pytest
import torch

from marl_optimal_execution_b200.qnet import DEFAULT_DIMS, QNetwork, init_params, torch_reference

pytestmark = pytest.mark.gpu


def states(n, seed, stride=8, offset=6, n_in=2):
    rs = np.random.RandomState(seed)
    x = rs.uniform(-5, 5, size=(n, stride))
    x[:, offset:offset + n_in] = rs.randint(0, 200, size=(n, n_in))       # digitised features: integers 0..199
    return torch.from_numpy(x).cuda()


def check(net, x, offset, tol=2e-4):
    q, act = net.forward(x, x_offset=offset)
    ref = torch_reference(net.params, x[:, offset:offset + net.dims[0]], net.dims)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    err = float((q - ref).abs().max())
    assert err <= tol * scale + 1e-5, (err, scale)
    ra = ref.argmax(dim=1).to(torch.int32)
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 4 * (tol * scale + 1e-5)
    assert bool((act[clear] == ra[clear]).all()) and float(clear.float().mean()) > 0.5
    assert bool((act == q.argmax(dim=1).to(torch.int32)).all())                # the kernel's own rule: first maximum of its Q row
    return err / max(scale, 1e-30)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 8192, 40000])
def test_reference_network_matches_torch_fp32(n):
    net = QNetwork(DEFAULT_DIMS, seed=11)
    rel = check(net, states(n, n), 6)
    assert rel < 2e-4


def test_trained_like_weights_biases_and_other_shapes():
    rs = np.random.RandomState(5)
    for dims in [(2, 32, 64, 128, 128, 64, 32, 24), (6, 32, 64, 128, 128, 64, 32, 24), (3, 16, 24), (16, 128, 128, 7), (1, 48, 24)]:
        p = init_params(dims, seed=3) * 3.0
        p += rs.normal(0, 0.05, size=p.shape).astype(np.float32)            # non-zero biases too
        net = QNetwork(dims, params=p)
        x = states(1000, 9, stride=dims[0] + 3, offset=2, n_in=dims[0])
        check(net, x, 2)


NNMODEL_2_DIMS = (2, 32, 64, 128, 256, 128, 64, 32, 24)      # util/model/QNets.py:30-52


@pytest.mark.parametrize("n", [1, 129, 9472])
def test_nnmodel_2_runs_through_the_streamed_kernel(n):
    """The 256-wide NNModel_2 (347 KB of hi + lo weights: does not fit in shared memory) through abx_qnet_forward_streamed_kernel -- weight blocks of 128 x 128 streamed per
    tile, two N blocks in separate TMEM column ranges for the 128 -> 256 layer, two accumulating K blocks for the 256 -> 128 layer -- against the PyTorch fp32 network."""
    rs = np.random.RandomState(21)
    p = init_params(NNMODEL_2_DIMS, seed=5) * 2.0
    p += rs.normal(0, 0.05, size=p.shape).astype(np.float32)
    net = QNetwork(NNMODEL_2_DIMS, params=p)
    rel = check(net, states(n, n + 3), 6)
    assert rel < 2e-4
    p2 = init_params(NNMODEL_2_DIMS, seed=6)                                # the learner's device-side weight push packs the same block layout
    net.set_params_device(torch.from_numpy(p2).cuda())
    net.params = p2
    check(net, states(n, n + 4), 6)


def test_streamed_kernel_equals_resident_kernel_on_the_default_network(monkeypatch):
    """ABX_QNET_FORCE_STREAMED=1 sends a network that fits through the streamed kernel: same Q rows as the resident kernel (same MMA sequence per layer)."""
    x = states(5000, 77)
    a = QNetwork(DEFAULT_DIMS, seed=4)
    qa, acta = a.forward(x, x_offset=6)
    monkeypatch.setenv("ABX_QNET_FORCE_STREAMED", "1")
    b = QNetwork(DEFAULT_DIMS, seed=4)
    qb, actb = b.forward(x, x_offset=6)
    torch.cuda.synchronize()
    assert torch.equal(qa, qb) and torch.equal(acta, actb)
    for dims in [(3, 16, 24), (16, 128, 128, 7), (8, 256, 256, 24), (4, 200, 136, 9)]:
        p = init_params(dims, seed=9) * 2.0
        net = QNetwork(dims, params=p)
        check(net, states(700, 3, stride=dims[0] + 3, offset=2, n_in=dims[0]), 2)


def test_set_params_and_epsilon_rule():
    net = QNetwork(DEFAULT_DIMS, seed=1)
    x = states(20000, 4)
    q0, a0 = net.forward(x, x_offset=6)
    net.set_params(init_params(DEFAULT_DIMS, seed=2))
    q1, a1 = net.forward(x, x_offset=6)
    assert not torch.equal(q0, q1)
    check(net, x, 6)
    # greedy_prob 0.9 (the reference's epsilon_max): ~10 % uniform actions, deterministic in (seed, counter)
    _, e1 = net.forward(x, x_offset=6, greedy_prob=0.9, seed=7, counter=3)
    _, e2 = net.forward(x, x_offset=6, greedy_prob=0.9, seed=7, counter=3)
    _, e3 = net.forward(x, x_offset=6, greedy_prob=0.9, seed=7, counter=4)
    assert torch.equal(e1, e2) and not torch.equal(e1, e3)
    frac = float((e1 != a1).float().mean())
    assert 0.06 < frac < 0.13, frac                                         # 0.1 * (1 - 1/24) expected
    _, r = net.forward(x, x_offset=6, greedy_prob=0.0, seed=1, counter=0)
    counts = torch.bincount(r.to(torch.int64), minlength=24).float()
    assert int(r.min()) >= 0 and int(r.max()) <= 23 and float(counts.min()) > 0.7 * 20000 / 24


def test_acting_and_learning_loop_on_the_gpu(golden_dir):
    """Closed loop: tcgen05 Q-network picks the actions from the environment's device-resident observation, the environment steps,
    the learner (ddqn.py) trains on the experience tuples and pushes new weights into the acting network."""
    import os
    from marl_optimal_execution_b200 import _lib
    from marl_optimal_execution_b200.ddqn import DDQNTrainer
    from marl_optimal_execution_b200.env import DDQNExecutionEnv
    g = np.load(os.path.join(golden_dir, "ddqn_IBM_2003-01-14_s4242.npz"))
    n = 512
    env = DDQNExecutionEnv(g["stream"], n_envs=n)
    env.reset(seeds=np.arange(n, dtype=np.uint64))
    tr = DDQNTrainer(device="cuda", batch_size=256, seed=2, buffer_capacity=1 << 16)
    net = QNetwork(DEFAULT_DIMS, params=tr.eval_net.flat())
    seen = []

    def act(obs, greedy_prob, tick):
        q, a = net.forward(obs, x_offset=6, greedy_prob=greedy_prob, seed=5, counter=tick)
        if tick == 20:
            ref = torch_reference(net.params, obs[:, 6:8])
            assert float((q - ref).abs().max()) <= 2e-4 * float(ref.abs().max()) + 1e-5
        seen.append(int(torch.unique(a).numel()))
        return a

    total, ticks = tr.run_episode(env, act, sync_fn=net.set_params, max_ticks=40)
    assert ticks == 40 and tr.learn_step_counter >= 6 and tr.buffer.size >= 30 * n
    assert max(seen) > 10 and float(total.min()) > 0.0 and np.isfinite([float(c) for c in tr.cost_hist]).all()
    st = env.stats()
    assert (st["flags"] & _lib.F_ERROR_MASK == 0).all()
    assert np.array_equal(net.params, tr.eval_net.flat())
    # device-side weight push (no host round trip) builds the same operand image as the host path
    x = states(3000, 1)
    q_host, _ = net.forward(x, x_offset=6)
    net.set_params(init_params(DEFAULT_DIMS, seed=77))
    net.set_params_device(tr.eval_net.flat_device())
    q_dev, _ = net.forward(x, x_offset=6)
    assert torch.equal(q_host, q_dev) and np.array_equal(net.sync_host_params(), tr.eval_net.flat())


def test_closed_loop_reproduces_a_reference_run_with_a_real_network(golden_dir):
    """End to end on the GPU: the tcgen05 Q-network picks every action from the environment's device-resident observation, the environment
    steps -- and the whole run equals a recorded reference run whose agent used the same weights (numpy fp32 stand-in for the Keras model,
    tools/record_reference_ddqn.py --mlp): same 660 actions, same 205 k events in the same order, same rewards."""
    import os
    from marl_optimal_execution_b200.env import DDQNExecutionEnv, dq_config
    g = np.load(os.path.join(golden_dir, "ddqn_mlp_IBM_2003-01-15_s31.npz"))
    stream = np.load(os.path.join(golden_dir, str(g["stream_fixture"])))["stream"]
    env = DDQNExecutionEnv(stream, n_envs=2, cfg=dq_config(hash_pops=1))
    env.reset(mom_sizes=np.tile(g["mom_sizes"].astype(np.int32), (2, 1)))
    net = QNetwork(DEFAULT_DIMS, params=g["mlp_params"])
    obs, trans, rew, done = env.step(torch.zeros(2, dtype=torch.int32, device="cuda"))
    acts, total, k = [], 0.0, 0
    ref_q = torch.from_numpy(g["mlp_q"]).cuda()
    while not bool(done[0]):
        q, a = net.forward(obs, x_offset=6)
        assert float((q[0] - ref_q[k]).abs().max()) <= 2e-4 * float(ref_q[k].abs().max()) + 1e-5, k
        acts.append(int(a[0]))
        obs, trans, rew, done = env.step(a)
        total += float(rew[0])
        k += 1
    assert acts == [int(x) for x in g["actions"]]
    st = env.stats()
    assert (st["messages"] == int(g["n_pops"])).all() and (st["pop_hash"] == np.uint64(int(g["pop_hash_ckpt"][-1]))).all() and (st["flags"] == 1).all()
    assert abs(total - float(g["step_reward_hist"].sum())) < 1e-9 * abs(total)
