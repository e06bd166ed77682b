"""oracle/sac_oracle.py — CPU restatement of continuous Soft Actor-Critic's `learn_batch`
(TEST INFRASTRUCTURE ONLY; eager PyTorch fp32 like the reference).

Restated reference sites (paths relative to /root/reference/pearl):
  policy_learners/sequential_decision_making/actor_critic_base.py:309-366     actor step, then critic step, soft target update
  policy_learners/sequential_decision_making/soft_actor_critic_continuous.py:131-231  losses, entropy autotune
  neural_networks/sequential_decision_making/actor_networks.py:29-51,488-591  GaussianActorNetwork, action_scaling
  neural_networks/sequential_decision_making/twin_critic.py:75-91, q_value_networks.py:152-174  twin VanillaQValueNetwork
  utils/functional_utils/learning/critic_utils.py:103-122,170-203             twin loss, target update
The two reparameterisation noise draws per step (Normal.rsample in sample_action: first on `state` for the actor
loss, then on `next_state` for the critic target) are INPUTS here, so that both sides of a parity test consume
identical noise.  Parity pinned by tests/golden/sac_small.npz (oracle/gen_golden.py).
"""
from __future__ import annotations

import math

import torch

from .pearl_oracle import _mlp, flat, load_flat  # noqa: F401


class _Actor(torch.nn.Module):
    def __init__(self, obs, act, hidden):
        super().__init__()
        dims = [obs] + list(hidden)
        self.body = torch.nn.Sequential(*[torch.nn.Sequential(torch.nn.Linear(dims[i], dims[i + 1]), torch.nn.ReLU())
                                          for i in range(len(dims) - 1)])
        self.fc_mu = torch.nn.Linear(hidden[-1], act)
        self.fc_std = torch.nn.Linear(hidden[-1], act)


class OracleSAC:
    def __init__(self, obs, act, actor_hidden, critic_hidden, low, high, *, actor_lr=1e-3, critic_lr=1e-3, gamma=0.99,
                 tau=0.005, entropy_coef=0.2, autotune=True, init=None):
        self.obs, self.act, self.gamma, self.tau, self.autotune = obs, act, gamma, tau, autotune
        self.low, self.high = torch.as_tensor(low, dtype=torch.float32), torch.as_tensor(high, dtype=torch.float32)
        self.bound = (self.high - self.low) / 2
        self.actor = _Actor(obs, act, actor_hidden)
        self.q = [_mlp([obs + act] + list(critic_hidden) + [1]) for _ in range(2)]
        self.qt = [_mlp([obs + act] + list(critic_hidden) + [1]) for _ in range(2)]
        if init is not None:
            load_flat(self.actor, init["actor"])
            for i in range(2):
                load_flat(self.q[i], init[f"q{i + 1}"])
                load_flat(self.qt[i], init[f"q{i + 1}t"])
        kw = dict(amsgrad=True)
        self.opt_actor = torch.optim.AdamW(self.actor.parameters(), lr=actor_lr, **kw)
        self.opt_critic = torch.optim.AdamW(list(self.q[0].parameters()) + list(self.q[1].parameters()), lr=critic_lr, **kw)
        self.log_alpha = torch.nn.Parameter(torch.zeros(1))
        self.opt_alpha = torch.optim.AdamW([self.log_alpha], lr=critic_lr, **kw)
        self.alpha = torch.exp(self.log_alpha).detach() if autotune else torch.tensor(entropy_coef)
        self.target_entropy = -torch.tensor(float(act))

    def sample_action(self, state, noise):
        x = self.actor.body(state)
        mean, z = self.actor.fc_mu(x), self.actor.fc_std(x)
        log_std = -5 + 0.5 * (2 - (-5)) * (torch.tanh(z) + 1)
        std = log_std.exp()
        sample = mean + std * noise                          # Normal(mean, std).rsample()
        na = torch.tanh(sample)
        action = (((self.high - self.low) * (na + 1.0)) / 2) + self.low
        log_prob = -((sample - mean) ** 2) / (2 * std ** 2) - log_std - math.log(math.sqrt(2 * math.pi))
        log_prob = log_prob - torch.log(self.bound * (1 - na.pow(2)) + 1e-6)
        return action, log_prob.sum(dim=1, keepdim=True)

    @staticmethod
    def _qv(net, s, a):
        return net(torch.cat([s, a], dim=-1)).squeeze(-1)

    def learn_batch(self, b, noise_actor, noise_next):
        s, a, r, s2, term = b["state"], b["action"], b["reward"], b["next_state"], b["terminated"]
        # ---- actor step
        act, logp = self.sample_action(s, noise_actor)
        q = torch.minimum(self._qv(self.q[0], s, act), self._qv(self.q[1], s, act)).unsqueeze(-1)
        actor_loss = (self.alpha * logp - q).mean()
        self.opt_actor.zero_grad()
        actor_loss.backward()
        self.opt_actor.step()
        # ---- critic step (with the UPDATED actor)
        self.opt_critic.zero_grad()
        with torch.no_grad():
            a2, logp2 = self.sample_action(s2, noise_next)
            nq = torch.minimum(self._qv(self.qt[0], s2, a2), self._qv(self.qt[1], s2, a2)).unsqueeze(-1)
            y = ((nq - self.alpha * logp2).view(-1) * self.gamma * (1 - term.float())) + r
        mse = torch.nn.MSELoss()
        critic_loss = (mse(self._qv(self.q[0], s, a), y) + mse(self._qv(self.q[1], s, a), y)) / 2.0
        critic_loss.backward()
        self.opt_critic.step()
        with torch.no_grad():
            for i in range(2):
                for pt, p in zip(self.qt[i].parameters(), self.q[i].parameters()):
                    pt.copy_(self.tau * p + (1.0 - self.tau) * pt)
        out = {"actor_loss": actor_loss.item(), "critic_loss": critic_loss.item()}
        if self.autotune:
            ent_loss = (-torch.exp(self.log_alpha) * (logp + self.target_entropy).detach()).mean()
            self.opt_alpha.zero_grad()
            ent_loss.backward()
            self.opt_alpha.step()
            self.alpha = torch.exp(self.log_alpha).detach()
            out["entropy_coef"] = ent_loss.item()
        return out
