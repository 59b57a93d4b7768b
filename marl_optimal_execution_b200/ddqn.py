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
        """-> (s, a, s', r, weight): np.random.choice(current_size, batch) (:463) over the valid rows.  Invalid rows carry a vanishing weight instead of zero
        so that a buffer holding only invalid rows cannot raise inside multinomial (that would take a device synchronisation to rule out beforehand); the
        returned weight is 1 for valid rows and 0 otherwise, and the caller's loss ignores the latter."""
        import torch
        idx = torch.multinomial(self.valid[: self.size] + 1e-30, batch, replacement=True, generator=generator)
        return self.s[idx], self.a[idx], self.sp[idx], self.r[idx], self.valid[idx]


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

    DROPOUT = 0.1      # util/model/QNets.py:16: Dropout(0.1) after every hidden layer but the first (:21-25); identity in predict(), active in train_on_batch()

    def __call__(self, x, training=False, generator=None):
        import torch
        h = x
        for i in range(0, len(self.params), 2):
            h = h @ self.params[i].T + self.params[i + 1]
            if i + 2 < len(self.params):
                h = torch.relu(h)
                if training and i >= 2 and self.DROPOUT > 0:             # Keras inverted dropout: kept units scaled by 1 / (1 - rate)
                    keep = (torch.rand(h.shape, device=h.device, generator=generator) >= self.DROPOUT).to(h.dtype)
                    h = h * keep / (1.0 - self.DROPOUT)
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
        self.rows_pushed, self.min_rows_per_push = 0, None              # the learn gate every rank evaluates identically (see agree_on_shard)

    def agree_on_shard(self, n_envs):
        """With several ranks training ONE policy every rank must take the learn / no-learn decision on the same ticks, or the gradient all-reduces pair up
        wrongly (or hang).  Shards may differ by one environment (distributed.shard_range), so the gate counts min-over-ranks rows per push: one tiny
        all-reduce here, none per step."""
        import torch
        import torch.distributed as dist
        m = int(n_envs)
        if self.sync_gradients and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            t = torch.tensor([m], dtype=torch.int64, device=self.device if str(self.device) != "cpu" else "cpu")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            m = int(t.item())
        self.min_rows_per_push = m
        return m

    def store(self, trans):
        """buffer.push + the rank-agreed row count behind the learn gate"""
        n = self.buffer.push(trans)
        self.rows_pushed += n if self.min_rows_per_push is None else min(n, self.min_rows_per_push)
        return n

    def can_learn(self):
        return min(self.rows_pushed, self.buffer.capacity) > self.batch_size                      # num_effective_experience > batch_size :262

    def learn(self):
        """One train_neural_nets call (:449-505) on a batch from the shared buffer."""
        import torch
        if not (self.can_learn() if self.rows_pushed else self.buffer.size > self.batch_size):       # num_effective_experience > batch_size :262
            return None
        s, a, sp, r, wt = self.buffer.sample(self.batch_size, self.gen)
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
        err = (self.eval_net(s, training=True, generator=self.gen) - q_target) ** 2                 # train_on_batch: Dropout(0.1) active (util/model/QNets.py:16-25); loss="mse" over all outputs
        loss = (err.mean(dim=1) * wt).sum() / wt.sum().clamp(min=1.0)                               # rows that hold no transition (weight 0) do not train
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
        return self.epsilon if (self.rows_pushed or self.buffer.size) + 1 > self.batch_size else 0.0

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
            self.store(torch.as_tensor(trans))
            total += torch.as_tensor(rew).to(self.device)
            if self.train_step_counter % self.train_every == 0 and self.learn() is not None and sync_fn is not None:
                sync_fn(self.eval_net.flat())
            self.train_step_counter += 1
            tick += 1
        return total, tick

    def run_episodes(self, env, act_fn, n_episodes, sync_fn=None, on_episode=None):
        """The reference's training sweep (config/execution/marketreplay/execution_marketreplay_ddqn_parallel.py:40-75: one simulation per train date, the
        network carried from one to the next) on the batched environment: `env` replays several days (environment e starts on day e % n_days), auto-reset
        moves every finished environment on to its next day, and one policy learns from all of them.  Returns the list of per-episode dicts
        (mean / std of the total step reward over environments, ticks, learn steps, mean loss)."""
        import torch
        env.set_auto_reset("next_day")
        self.agree_on_shard(env.n_envs)
        hist = []
        for ep in range(n_episodes):
            l0, c0 = self.learn_step_counter, len(self.cost_hist)
            total, ticks = self.run_episode(env, act_fn, sync_fn)
            costs = torch.stack(self.cost_hist[c0:]).float() if len(self.cost_hist) > c0 else torch.zeros(1)
            rec = {"episode": ep, "ticks": ticks, "mean_total_reward": float(total.mean()), "std_total_reward": float(total.std()), "learn_steps": self.learn_step_counter - l0,
                   "mean_loss": float(costs.mean()), "epsilon": float(self.epsilon)}
            hist.append(rec)
            if on_episode is not None:
                on_episode(rec)
        return hist
