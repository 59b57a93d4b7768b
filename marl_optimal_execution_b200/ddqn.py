"""DDQN learner over the batched execution environment.

The reference trains one Keras network inside one agent of one simulation (agent/execution/qlearning/
ddqlearning_execution_agent.py:449-505, train_neural_nets) and repeats whole simulations per day / per process
(config/execution/marketreplay/execution_marketreplay_ddqn_parallel.py).  Here one network acts in thousands of
environments at once (the tcgen05 forward of qnet.py picks the actions, DDQNExecutionEnv steps the books) and learns from
all of their experience tuples.  The update is the reference's, restated for a batch drawn from a shared replay buffer:

    q_next      = target(s')                      :470
    q_eval4next = target(s')                      :472  (the reference asks the TARGET network twice: its "double" DQN
                                                         selects and evaluates the next action with the same network)
    q_target    = eval(s);  q_target[a] = r + gamma * q_next[argmax q_eval4next]     :477-485
    every `replace_target_iter` learn steps: target <- eval (before the fit)         :487-490
    eval.train_on_batch(s, q_target): MSE over all outputs, RMSprop(lr)              :492-494, :111-115
    epsilon += epsilon_increment up to epsilon_max                                   :503

PyTorch holds the buffers and does the fp32 autograd of the 38 k-parameter MLP (plumbing around the simulator, which is the
product); the acting path never leaves the GPU.  `act_fn` lets the CPU test-suite drive the same loop with a host policy.
"""
import numpy as np

from .qnet import DEFAULT_DIMS, init_params, param_count, unpack_params


class ReplayBuffer:
    """Ring of finalised experience tuples (s[2], a, s'[2], r) on one device; the reference keeps an OrderedDict per agent (:119).
    No operation here synchronises the device: every push writes a fixed number of rows with a validity flag, sampling draws among valid rows."""

    def __init__(self, capacity, device):
        import torch
        self.capacity, self.device = int(capacity), device
        self.s = torch.zeros(self.capacity, 2, dtype=torch.float32, device=device)
        self.sp = torch.zeros(self.capacity, 2, dtype=torch.float32, device=device)
        self.a = torch.zeros(self.capacity, dtype=torch.int64, device=device)
        self.r = torch.zeros(self.capacity, dtype=torch.float32, device=device)
        self.valid = torch.zeros(self.capacity, dtype=torch.float32, device=device)
        self.size, self.head = 0, 0

    def push(self, trans):
        """trans: [n, 6] (s0, s1, a, s'0, s'1, r) as DDQNExecutionEnv.step returns it.  Rows with r NaN (the reference's None: neither accepted
        nor executed) or without a transition are stored as invalid and never sampled: they would break np.array arithmetic in :485."""
        import torch
        t = trans.to(self.device)
        n = int(t.shape[0])
        if n == 0:
            return 0
        if n > self.capacity:
            t, n = t[-self.capacity:], self.capacity
        ok = ~torch.isnan(t).any(dim=1)
        t = torch.nan_to_num(t, nan=0.0)
        idx = (self.head + torch.arange(n, device=self.device)) % self.capacity
        self.s[idx] = t[:, 0:2].float(); self.a[idx] = t[:, 2].long(); self.sp[idx] = t[:, 3:5].float(); self.r[idx] = t[:, 5].float(); self.valid[idx] = ok.float()
        self.head = (self.head + n) % self.capacity
        self.size = min(self.size + n, self.capacity)
        return n

    def n_valid(self):
        return int(self.valid[: self.size].sum()) if self.size else 0

    def sample(self, batch, generator=None):
        import torch
        idx = torch.multinomial(self.valid[: self.size], batch, replacement=True, generator=generator)   # np.random.choice(current_size, batch) :463 over the valid rows
        return self.s[idx], self.a[idx], self.sp[idx], self.r[idx]


class TorchMLP:
    """fp32 autograd copy of the Q-network (util/model/QNets.py:7-27) for the learner; parameters in the flat layout of qnet.py."""

    def __init__(self, dims=DEFAULT_DIMS, flat=None, device="cpu", seed=0):
        import torch
        self.dims = tuple(dims)
        flat = init_params(self.dims, seed) if flat is None else np.asarray(flat, dtype=np.float32)
        assert flat.size == param_count(self.dims)
        self.params = []
        for w, b in unpack_params(flat, self.dims):
            self.params += [torch.tensor(w, device=device, requires_grad=True), torch.tensor(b, device=device, requires_grad=True)]

    def __call__(self, x):
        import torch
        h = x
        for i in range(0, len(self.params), 2):
            h = h @ self.params[i].T + self.params[i + 1]
            if i + 2 < len(self.params):
                h = torch.relu(h)
        return h

    def flat(self):
        return self.flat_device().cpu().numpy().astype(np.float32)

    def flat_device(self):
        import torch
        with torch.no_grad():
            return torch.cat([p.reshape(-1) for p in self.params]).contiguous()

    def load(self, other):
        import torch
        with torch.no_grad():
            for p, q in zip(self.params, other.params):
                p.copy_(q)


class DDQNTrainer:
    def __init__(self, dims=DEFAULT_DIMS, device="cpu", batch_size=32, learning_rate=0.01, reward_decay=0.98, replace_target_iter=5,
                 epsilon_max=0.9, epsilon_increment=None, train_every=5, buffer_capacity=1 << 20, seed=0, sync_gradients=True):
        """Defaults are the agent's constructor defaults (:40-66)."""
        import torch
        self.device, self.batch_size, self.gamma = device, int(batch_size), float(reward_decay)
        self.replace_target_iter, self.train_every = int(replace_target_iter), int(train_every)
        self.epsilon_max, self.epsilon_increment = float(epsilon_max), epsilon_increment
        self.epsilon = 0.0 if epsilon_increment is not None else self.epsilon_max                     # :100
        self.eval_net = TorchMLP(dims, device=device, seed=seed)
        self.target_net = TorchMLP(dims, device=device, seed=seed + 1)
        self.opt = torch.optim.RMSprop(self.eval_net.params, lr=learning_rate, alpha=0.9, eps=1e-7)   # Keras RMSprop defaults (rho 0.9, epsilon 1e-7)
        self.buffer = ReplayBuffer(buffer_capacity, device)
        self.gen = torch.Generator(device=device); self.gen.manual_seed(seed)
        self.learn_step_counter, self.train_step_counter, self.cost_hist = 0, 0, []
        self.sync_gradients = bool(sync_gradients)

    def learn(self):
        """One train_neural_nets call (:449-505) on a batch from the shared buffer."""
        import torch
        if self.buffer.size <= self.batch_size:                                                     # num_effective_experience > batch_size :262
            return None
        s, a, sp, r = self.buffer.sample(self.batch_size, self.gen)
        with torch.no_grad():
            q_next = self.target_net(sp)
            q_eval4next = self.target_net(sp)
            q_target = self.eval_net(s).clone()
            max_act4next = q_eval4next.argmax(dim=1)
            sel = q_next.gather(1, max_act4next[:, None])[:, 0]
            q_target[torch.arange(self.batch_size, device=s.device), a] = r + self.gamma * sel
        if self.learn_step_counter % self.replace_target_iter == 0:
            self.target_net.load(self.eval_net)
        self.opt.zero_grad(set_to_none=True)
        loss = ((self.eval_net(s) - q_target) ** 2).mean()                                          # loss="mse" over all outputs
        loss.backward()
        self.allreduce_gradients()
        self.opt.step()
        cost = loss.detach()                                                                       # stays on the device: no synchronisation per update
        self.cost_hist.append(cost)
        if self.epsilon_increment is not None:
            self.epsilon = self.epsilon + self.epsilon_increment if self.epsilon < self.epsilon_max else self.epsilon_max
        self.learn_step_counter += 1
        return cost

    def allreduce_gradients(self):
        """One policy trained by all ranks: average the 38 k fp32 gradients over the process group (NCCL over NVLink on the
        GPU box, gloo in the CPU tests) before the optimizer step; every rank then applies the same update to identical weights.  No-op for a
        single process or with sync_gradients=False (independent learners)."""
        import torch
        import torch.distributed as dist
        if not self.sync_gradients or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in self.eval_net.params])
        dist.all_reduce(flat)
        flat /= dist.get_world_size()
        off = 0
        for p in self.eval_net.params:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n

    def greedy_prob(self):
        """Probability of the network's action in choose_action (:349-357): epsilon once the buffer can feed a batch, else 0 (all random)."""
        return self.epsilon if self.buffer.size + 1 > self.batch_size else 0.0

    def run_episode(self, env, act_fn, sync_fn=None, max_ticks=None):
        """One pass over the batched environment: act, step, store, learn every `train_every` ticks (:259-266).
        act_fn(obs, greedy_prob, tick) -> int32 actions [n_envs]; sync_fn(flat_params) is called after each learn step (push the new
        weights to the acting network).  The caller has reset `env`.  Returns (total reward per environment, ticks)."""
        import torch
        first = None if str(self.device) == "cpu" else torch.zeros(env.n_envs, dtype=torch.int32, device=self.device)   # ignored: no decision is pending yet
        obs, trans, rew, done = env.step(first)
        total = torch.zeros(env.n_envs, dtype=torch.float64, device=self.device)
        tick = 0
        while True:
            d = torch.as_tensor(done)
            if bool(d.all()) or (max_ticks is not None and tick >= max_ticks):
                break
            actions = act_fn(obs, self.greedy_prob(), tick)
            obs, trans, rew, done = env.step(actions)
            self.buffer.push(torch.as_tensor(trans))
            total += torch.as_tensor(rew).to(self.device)
            if self.train_step_counter % self.train_every == 0 and self.learn() is not None and sync_fn is not None:
                sync_fn(self.eval_net.flat())
            self.train_step_counter += 1
            tick += 1
        return total, tick
