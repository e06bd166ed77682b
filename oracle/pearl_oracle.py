"""oracle/pearl_oracle.py — CPU restatement of Pearl's learner hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
`cpu_baseline` / `--impl reference` legs may import this module; the product
(`pearl_b200/`) never does and fails loudly without its CUDA library.

The reference is eager PyTorch fp32 on CPU, so this restatement is eager
PyTorch fp32 on CPU too (same ATen ops in the same order => bit-identical to
the reference on the machine that generated tests/golden/*.npz; elsewhere the
tests allow 2e-6).  Index selection goes through CPython's `random.sample`
exactly like the reference; `oracle/mt_sample_oracle.c` restates that part in
C and is pinned by tests/golden/random_sample_kat.json.

Parity is PINNED: tests/test_oracle_golden.py replays every fixture recorded by
oracle/gen_golden.py from the real reference (`/root/reference`, commit
48f1fbb) and requires identical indices and q / y / parameters / AdamW state.

Reference sites restated here (paths relative to /root/reference/pearl):
  replay_buffers/tensor_based_replay_buffer.py:55-133,179-251  push, padding+mask
  replay_buffers/basic_replay_buffer.py:21-48                 deque(maxlen) FIFO
  replay_buffers/tensor_based_replay_buffer.py:253-400        sample + collate
  policy_learners/policy_learner.py:162-218                   learn loop, preprocess
  action_representation_modules/one_hot_...module.py:27-34    one-hot
  neural_networks/sequential_decision_making/q_value_networks.py:152-174
  neural_networks/common/utils.py:75-152,214-226              mlp_block, soft update
  policy_learners/sequential_decision_making/deep_td_learning.py:269-360
  .../deep_q_learning.py:130-167, .../double_dqn.py:29-57     bootstrap targets
  torch/optim/adam.py (AdamW, amsgrad=True, wd=0.01)          third-party
"""
from __future__ import annotations

import random
from collections import deque

import torch
import torch.nn.functional as F


class OracleReplayBuffer:
    """FIFO deque of per-transition tensor tuples, sampled with `random.sample`."""

    def __init__(self, capacity: int, n_actions: int) -> None:
        self.capacity = capacity
        self.n_actions = n_actions
        self.memory: deque = deque([], maxlen=capacity)

    def __len__(self) -> int:
        return len(self.memory)

    def clear(self) -> None:
        self.memory = deque([], maxlen=self.capacity)

    def push(self, state, action, reward, terminated, truncated, next_state,
             next_available_ids=None) -> None:
        A = self.n_actions
        ids = list(range(A)) if next_available_ids is None else [int(a) for a in next_available_ids]
        # padded [A,1] float action tensor + bool mask (True = unavailable), :179-251
        avail = torch.zeros((A, 1), dtype=torch.float32)
        avail[: len(ids), 0] = torch.tensor(ids, dtype=torch.float32)
        mask = torch.zeros((A,))
        mask[len(ids):] = 1
        self.memory.append((
            torch.as_tensor(state).clone().unsqueeze(0),
            torch.tensor(int(action)).unsqueeze(0),
            torch.tensor([reward]),
            torch.tensor([bool(terminated)]),
            torch.tensor([bool(truncated)]),
            torch.as_tensor(next_state).clone().unsqueeze(0),
            avail.unsqueeze(0),
            mask.bool().unsqueeze(0),
        ))

    def sample(self, batch_size: int) -> dict:
        if batch_size > len(self):
            raise ValueError(
                f"Can't get a batch of size {batch_size} from a replay buffer with "
                f"only {len(self)} elements")
        picked = random.sample(self.memory, batch_size)
        cols = list(zip(*picked))
        return dict(
            state=torch.cat(cols[0]).type(torch.float32),
            action=torch.cat(cols[1]),
            reward=torch.cat(cols[2]),
            terminated=torch.cat(cols[3]),
            truncated=torch.cat(cols[4]),
            next_state=torch.cat(cols[5]).type(torch.float32),
            next_available_actions=torch.cat(cols[6]),
            next_unavailable_actions_mask=torch.cat(cols[7]),
        )


def _mlp(dims):
    layers = []
    for i in range(len(dims) - 2):
        layers.append(torch.nn.Sequential(torch.nn.Linear(dims[i], dims[i + 1]), torch.nn.ReLU()))
    layers.append(torch.nn.Sequential(torch.nn.Linear(dims[-2], dims[-1])))
    return torch.nn.Sequential(*layers)


def flat(module) -> torch.Tensor:
    return torch.cat([p.detach().reshape(-1) for p in module.parameters()])


def load_flat(module, vec) -> None:
    vec = torch.as_tensor(vec, dtype=torch.float32)
    off = 0
    with torch.no_grad():
        for p in module.parameters():
            n = p.numel()
            p.copy_(vec[off:off + n].view_as(p))
            off += n
    assert off == vec.numel()


class OracleDQN:
    """DeepQLearning / DoubleDQN `learn()` as the reference computes it."""

    def __init__(self, obs, n_actions, hidden, *, lr=1e-3, gamma=0.99, batch_size=128,
                 training_rounds=10, target_update_freq=10, tau=0.75, double=False,
                 weight_decay=0.01, init_q=None, init_q_target=None) -> None:
        self.obs, self.A = obs, n_actions
        self.gamma, self.tau = gamma, tau
        self.batch_size, self.training_rounds = batch_size, training_rounds
        self.target_update_freq, self.double = target_update_freq, double
        self.Q = _mlp([obs + n_actions] + list(hidden) + [1])
        self.Qt = _mlp([obs + n_actions] + list(hidden) + [1])
        if init_q is not None:
            load_flat(self.Q, init_q)
        load_flat(self.Qt, init_q_target if init_q_target is not None else flat(self.Q))
        self.opt = torch.optim.AdamW(self.Q.parameters(), lr=lr, amsgrad=True,
                                     weight_decay=weight_decay)
        self.training_steps = 0
        self.trace = None  # optional dict of lists: idx-free q / y per round

    # q_value_networks.py:152-174 — state repeated per query action, action LAST
    def _q_values(self, net, state, action):
        act3 = action.unsqueeze(1) if action.dim() == 2 else action
        s = torch.repeat_interleave(state.unsqueeze(1), act3.shape[-2], dim=1)
        q = net(torch.cat([s, act3], dim=-1)).squeeze(-1)
        return q if action.dim() == 3 else q.squeeze(-1)

    def _one_hot(self, x):
        if x.dim() == 1:
            x = x.unsqueeze(-1)
        return F.one_hot(x.long(), num_classes=self.A).squeeze(dim=-2).float()

    @torch.no_grad()
    def _next_values(self, b):
        nxt, mask = b["next_available_actions"], b["next_unavailable_actions_mask"]
        if not self.double:                                      # deep_q_learning.py:130-167
            v = self._q_values(self.Qt, b["next_state"], nxt)
            v[mask] = -float("inf")
            return v.max(1)[0]
        v = self._q_values(self.Q, b["next_state"], nxt)        # double_dqn.py:29-57
        v[mask] = -float("inf")
        a_star = v.max(1)[1]
        chosen = nxt[torch.arange(nxt.size(0)), a_star.squeeze()]
        return self._q_values(self.Qt, b["next_state"], chosen)

    def learn_batch(self, b) -> float:
        if (self.training_steps + 1) % self.target_update_freq == 0:   # deep_td_learning.py:283-284
            with torch.no_grad():
                for pt, p in zip(self.Qt.parameters(), self.Q.parameters()):
                    pt.copy_(self.tau * p + (1.0 - self.tau) * pt)
        q = self._q_values(self.Q, b["state"], b["action"])
        y = self._next_values(b) * self.gamma * (1 - b["terminated"].float()) + b["reward"]
        if b.get("weight") is not None:   # prioritized replay: importance-weighted squared TD error
            loss = (b["weight"] * (q - y) ** 2).mean()
        else:
            loss = torch.nn.MSELoss()(q, y)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        if self.trace is not None:
            self.trace["q"].append(q.detach().clone())
            self.trace["y"].append(y.detach().clone())
        return torch.abs(q - y).mean().item()

    def learn(self, buf: OracleReplayBuffer) -> dict:
        if len(buf) == 0:
            return {}
        bs = len(buf) if (self.batch_size == -1 or len(buf) < self.batch_size) else self.batch_size
        report = {"loss": []}
        for _ in range(self.training_rounds):
            self.training_steps += 1
            b = buf.sample(bs)
            b["action"] = self._one_hot(b["action"])                       # preprocess_batch
            b["next_available_actions"] = self._one_hot(b["next_available_actions"])
            report["loss"].append(self.learn_batch(b))
        return report
