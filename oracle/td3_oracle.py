"""oracle/td3_oracle.py — CPU restatement of the `learn_batch` of Pearl's deterministic actor-critic learners
(TEST INFRASTRUCTURE ONLY; eager PyTorch fp32 like the reference).

Restated reference sites (paths relative to /root/reference/pearl):
  policy_learners/sequential_decision_making/ddpg.py:105-157        DeepDeterministicPolicyGradient actor / critic losses
  policy_learners/sequential_decision_making/td3.py:106-202         TD3: delayed actor + target updates, clipped target noise
  policy_learners/sequential_decision_making/actor_critic_base.py:309-366   actor step, critic step, target updates (DDPG)
  neural_networks/sequential_decision_making/actor_networks.py:29-51,448-485  VanillaContinuousActorNetwork (tanh), action_scaling
  neural_networks/sequential_decision_making/twin_critic.py:75-91, utils/functional_utils/learning/critic_utils.py:103-122,170-203
DDPG in this reference trains a TWIN critic as well (critic loss = (mse1 + mse2) / 2, target = min of the target twins);
its actor loss uses critic 1 only.  TD3 = DDPG + `actor_update_freq` (actor and ALL target updates only when
`training_steps % freq == 0`) + clipped Gaussian noise on the target action (`torch.normal`, scaled by (high - low) / 2,
clamped to the box).  The noise draws are INPUTS here so that both sides of a parity test consume identical noise.
Parity pinned by tests/golden/td3_small.npz and ddpg_small.npz (oracle/gen_golden.py).
"""
from __future__ import annotations

import torch

from .pearl_oracle import _mlp, flat, load_flat  # noqa: F401


def _actor(obs, act, hidden):
    dims = [obs] + list(hidden)
    layers = [torch.nn.Sequential(torch.nn.Linear(dims[i], dims[i + 1]), torch.nn.ReLU()) for i in range(len(dims) - 1)]
    layers.append(torch.nn.Sequential(torch.nn.Linear(dims[-1], act), torch.nn.Tanh()))
    return torch.nn.Sequential(*layers)


class OracleTD3:
    def __init__(self, obs, act, actor_hidden, critic_hidden, low, high, *, actor_lr=1e-3, critic_lr=1e-3, gamma=0.99,
                 actor_tau=0.005, critic_tau=0.005, actor_update_freq=2, noise_clip=0.5, init=None):
        self.gamma, self.actor_tau, self.critic_tau, self.freq, self.noise_clip = gamma, actor_tau, critic_tau, actor_update_freq, noise_clip
        self.low, self.high = torch.as_tensor(low, dtype=torch.float32), torch.as_tensor(high, dtype=torch.float32)
        self.actor, self.actor_t = _actor(obs, act, actor_hidden), _actor(obs, act, actor_hidden)
        self.q = [_mlp([obs + act] + list(critic_hidden) + [1]) for _ in range(2)]
        self.qt = [_mlp([obs + act] + list(critic_hidden) + [1]) for _ in range(2)]
        if init is not None:
            load_flat(self.actor, init["actor"]); load_flat(self.actor_t, init["actor_t"])
            for i in range(2):
                load_flat(self.q[i], init[f"q{i + 1}"]); load_flat(self.qt[i], init[f"q{i + 1}t"])
        self.opt_actor = torch.optim.AdamW(self.actor.parameters(), lr=actor_lr, amsgrad=True)
        self.opt_critic = torch.optim.AdamW(list(self.q[0].parameters()) + list(self.q[1].parameters()), lr=critic_lr, amsgrad=True)
        self.training_steps = 0
        self.last_actor_loss = 0.0

    def act(self, net, s):
        return (((self.high - self.low) * (net(s) + 1.0)) / 2) + self.low

    @staticmethod
    def _qv(net, s, a):
        return net(torch.cat([s, a], dim=-1)).squeeze(-1)

    def learn_batch(self, b, target_noise=None):
        """One `learn_batch` (the caller advances `training_steps` first, as PolicyLearner.learn does).
        `target_noise`: the `torch.normal(0, actor_update_noise, ...)` draw of this step, or None for DDPG."""
        s, a, r, s2, term = b["state"], b["action"], b["reward"], b["next_state"], b["terminated"]
        update_actor = self.freq <= 1 or self.training_steps % self.freq == 0
        if update_actor:
            self.opt_actor.zero_grad()
            actor_loss = -self._qv(self.q[0], s, self.act(self.actor, s)).mean()
            actor_loss.backward()
            self.opt_actor.step()
            self.last_actor_loss = actor_loss.item()
        self.opt_critic.zero_grad()
        with torch.no_grad():
            a2 = self.act(self.actor_t, s2)
            if target_noise is not None:
                noise = torch.clamp(target_noise, -self.noise_clip, self.noise_clip) * (self.high - self.low) / 2
                a2 = torch.clamp(a2 + noise, self.low, self.high)
            nq = torch.minimum(self._qv(self.qt[0], s2, a2), self._qv(self.qt[1], s2, a2))
            y = (nq * self.gamma * (1 - term.float())) + r
        mse = torch.nn.MSELoss()
        critic_loss = (mse(self._qv(self.q[0], s, a), y) + mse(self._qv(self.q[1], s, a), y)) / 2.0
        critic_loss.backward()
        self.opt_critic.step()
        if update_actor:
            with torch.no_grad():
                for i in range(2):
                    for pt, p in zip(self.qt[i].parameters(), self.q[i].parameters()):
                        pt.copy_(self.critic_tau * p + (1.0 - self.critic_tau) * pt)
                for pt, p in zip(self.actor_t.parameters(), self.actor.parameters()):
                    pt.copy_(self.actor_tau * p + (1.0 - self.actor_tau) * pt)
        return {"actor_loss": self.last_actor_loss, "critic_loss": critic_loss.item()}
