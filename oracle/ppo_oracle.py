"""oracle/ppo_oracle.py — CPU restatement of PPO (TEST INFRASTRUCTURE ONLY; eager PyTorch fp32 like the reference).

`gae_reference_loop` follows pearl/policy_learners/sequential_decision_making/ppo.py:271-293 literally
(newest -> oldest, same tensor expressions), on arrays in time order.  Pinned by tests/golden/ppo_gae.npz,
recorded from the real reference by oracle/gen_golden.py, and by the closed form of the reference's own unit
test (test/unit/with_pytorch/test_ppo.py:48-115).

`OraclePPO` restates ProximalPolicyOptimization.learn: preprocess_replay_buffer (ppo.py:201-293: state values,
action probabilities of the taken actions under the current policy, GAE / lambda returns over the whole buffer),
then PolicyLearner.learn (policy_learner.py:162-204) sampling with `random.sample` and ActorCriticBase.learn_batch
(actor_critic_base.py:309-349): clipped-surrogate actor step (ppo.py:152-184, VanillaActorNetwork softmax policy,
actor_networks.py:107-176) followed by the state-value critic step (critic_utils.py:139-167).  Pinned by
tests/golden/ppo_small.npz.
"""
from __future__ import annotations

import random

import torch

from .pearl_oracle import _mlp, flat, load_flat  # noqa: F401


def gae_reference_loop(values, last_next_value, reward, terminated, truncated, discount_factor, trace_decay_param,
                       incoming_gae=0.0):
    """`incoming_gae`: the chain entering from newer transitions held elsewhere (0.0 = the reference's start value);
    used to check the sharded rollout protocol (pearl_b200/dist.py: sharded_gae_fixup)."""
    n = values.shape[0]
    gae_out = torch.empty(n, dtype=torch.float32)
    lam_out = torch.empty(n, dtype=torch.float32)
    next_value = torch.as_tensor(last_next_value, dtype=torch.float32).reshape(1)
    gae = torch.tensor([float(incoming_gae)])
    for t in range(n - 1, -1, -1):
        term, trunc = terminated[t].reshape(1), truncated[t].reshape(1)
        td_error = reward[t].reshape(1) + discount_factor * next_value * (~term) - values[t].reshape(1)
        gae = td_error + discount_factor * trace_decay_param * (not (bool(term) or bool(trunc))) * gae
        gae_out[t] = gae[0]
        lam_out[t] = (gae + values[t].reshape(1))[0]
        next_value = values[t].reshape(1)
    return gae_out, lam_out


class OraclePPO:
    def __init__(self, obs, n_actions, actor_hidden, critic_hidden, *, actor_lr=1e-4, critic_lr=1e-4, gamma=0.99, epsilon=0.0,
                 trace_decay=0.95, entropy_bonus=0.01, batch_size=128, training_rounds=100, init_actor=None, init_critic=None):
        self.obs, self.A, self.gamma, self.eps, self.lam, self.beta = obs, n_actions, gamma, epsilon, trace_decay, entropy_bonus
        self.batch_size, self.training_rounds = batch_size, training_rounds
        self.actor = _mlp([obs] + list(actor_hidden) + [n_actions])      # softmax applied in `probs`
        self.critic = _mlp([obs] + list(critic_hidden) + [1])
        if init_actor is not None:
            load_flat(self.actor, init_actor)
        if init_critic is not None:
            load_flat(self.critic, init_critic)
        self.opt_actor = torch.optim.AdamW(self.actor.parameters(), lr=actor_lr, amsgrad=True)
        self.opt_critic = torch.optim.AdamW(self.critic.parameters(), lr=critic_lr, amsgrad=True)

    def probs(self, state):
        return torch.softmax(self.actor(state), dim=-1)

    def action_prob(self, state, action_ids):
        onehot = torch.nn.functional.one_hot(action_ids.long(), self.A).float()
        return torch.sum(self.probs(state) * onehot, dim=1, keepdim=True).view(-1)

    @torch.no_grad()
    def preprocess(self, state, action, reward, terminated, truncated, last_next_state):
        """Arrays in time order (oldest first).  The reference evaluates the networks on the buffer newest-first; the
        per-row results do not depend on that order beyond GEMM blocking."""
        rev = torch.arange(state.shape[0] - 1, -1, -1)
        values = self.critic(state[rev]).reshape(-1)[rev]
        ap = self.action_prob(state[rev], action[rev])[rev]
        next_value = self.critic(last_next_state.reshape(1, -1)).reshape(-1)[0]
        gae, lam_return = gae_reference_loop(values, next_value, reward, terminated, truncated, self.gamma, self.lam)
        return dict(values=values, action_probs=ap, gae=gae, lam_return=lam_return)

    def learn_batch(self, state, action, gae, lam_return, action_probs_old):
        ap = self.action_prob(state, action)
        r = torch.div(ap, action_probs_old)
        clip = torch.clamp(r, min=1.0 - self.eps, max=1.0 + self.eps)
        loss = torch.sum(-torch.min(r * gae, clip * gae))
        entropy = torch.distributions.Categorical(ap.detach()).entropy()
        loss = loss - torch.sum(self.beta * entropy)
        self.opt_actor.zero_grad()
        loss.backward()
        self.opt_actor.step()
        self.opt_critic.zero_grad()
        vs = self.critic(state)
        closs = torch.nn.MSELoss()(vs.reshape_as(lam_return), lam_return.detach())
        closs.backward()
        self.opt_critic.step()
        return {"actor_loss": loss.item(), "critic_loss": closs.item()}

    def learn(self, state, action, reward, terminated, truncated, last_next_state, trace=None):
        n = state.shape[0]
        pre = self.preprocess(state, action, reward, terminated, truncated, last_next_state)
        B = n if (self.batch_size == -1 or n < self.batch_size) else self.batch_size
        report = {"actor_loss": [], "critic_loss": []}
        for _ in range(self.training_rounds):
            idx = torch.tensor(random.sample(range(n), B))
            if trace is not None:
                trace.setdefault("idx", []).append(idx.tolist())
            out = self.learn_batch(state[idx], action[idx], pre["gae"][idx], pre["lam_return"][idx], pre["action_probs"][idx])
            for k, v in out.items():
                report[k].append(v)
        return report, pre
