"""PPO learner on the GPU against the recording of the reference (tests/golden/ppo_small.npz) and against
oracle/ppo_oracle.py on a BASELINE configs[3]-shaped rollout slice (obs 210, [256, 256] networks)."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.pearl_oracle import flat
from oracle.ppo_oracle import OraclePPO

pytestmark = pytest.mark.gpu


from _tol import close as _close, close_params as _close_params  # elementwise 1e-4 (+ counted AdamW outliers)


def _rollout_buffer(states, action, reward, terminated, truncated, n_act, capacity=None, prefill=0):
    import pearl_b200
    n = action.shape[0]
    buf = pearl_b200.B200ReplayBuffer(capacity or n)
    t = torch.from_numpy
    if prefill:      # older transitions that the ring evicts again: the rollout then wraps around the ring end
        z = np.zeros((prefill, states.shape[1]), dtype=np.float32)
        buf.push_batch(t(z), torch.zeros(prefill, dtype=torch.int64), torch.zeros(prefill), t(z), torch.zeros(prefill, dtype=torch.bool),
                       torch.zeros(prefill, dtype=torch.bool), max_number_actions=n_act)
    buf.push_batch(t(states[:n]), t(action), t(reward), t(states[1:n + 1]), t(terminated), t(truncated), max_number_actions=n_act)
    return buf


@pytest.mark.parametrize("graph", [True, False])
def test_ppo_matches_reference_recording(graph):
    import pearl_b200
    fx = np.load(os.path.join(GOLDEN, "ppo_small.npz"))
    n, R, B, A = int(fx["n"]), int(fx["rounds"]), int(fx["batch"]), int(fx["n_act"])
    buf = _rollout_buffer(fx["states"], fx["action"], fx["reward"], fx["terminated"], fx["truncated"], A,
                          capacity=n, prefill=0 if graph else 150)
    pl = pearl_b200.B200ProximalPolicyOptimization(
        state_dim=int(fx["obs"]), n_actions=A, actor_hidden_dims=[32, 32], critic_hidden_dims=[32, 32], training_rounds=R, batch_size=B,
        actor_learning_rate=float(fx["actor_lr"]), critic_learning_rate=float(fx["critic_lr"]), discount_factor=float(fx["gamma"]),
        epsilon=float(fx["epsilon"]), trace_decay_param=float(fx["lam"]), entropy_bonus_scaling=float(fx["beta"]))
    pl.load_parameters(fx["init_actor"], fx["init_critic"])
    pl.use_cuda_graph = graph
    random.seed(53)
    trace = {}
    rep = pl.learn(buf, trace=trace)
    assert trace["idx"].tolist() == fx["idx"].tolist()
    pre = pl.last_preprocess
    _close(pre["action_probs"].cpu().numpy(), fx["action_probs"], "action_probs", rtol=2e-5)
    _close(pre["gae"].cpu().numpy(), fx["gae"], "gae", rtol=2e-5)
    _close(pre["lam_return"].cpu().numpy(), fx["lam_return"], "lam_return", rtol=2e-5)
    _close(rep["actor_loss"], fx["actor_loss"], "actor_loss")
    _close(rep["critic_loss"], fx["critic_loss"], "critic_loss")
    _close(pl.actor_params.cpu().numpy(), fx["actor_after"], "actor")
    _close(pl.critic_params.cpu().numpy(), fx["critic_after"], "critic")


def test_ppo_rollout_shape_against_oracle():
    """obs 210, [256, 256] networks (BASELINE configs[3]); a 20k-step rollout (three preprocessing chunks), batch 256."""
    import pearl_b200
    torch.manual_seed(9)
    torch.set_num_threads(4)
    obs, A, n, B, R = 210, 8, 20000, 256, 4
    rng = np.random.Generator(np.random.PCG64(21))
    q8 = lambda x: (np.rint(x * 256) / 256).astype(np.float32)
    states, reward = q8(rng.standard_normal((n + 1, obs))), q8(rng.standard_normal(n))
    action = rng.integers(0, A, size=n).astype(np.int64)
    terminated, truncated = rng.random(n) < 0.004, rng.random(n) < 0.002
    orc = OraclePPO(obs, A, (256, 256), (256, 256), actor_lr=3e-4, critic_lr=3e-4, gamma=0.99, epsilon=0.2, trace_decay=0.95,
                    entropy_bonus=0.01, batch_size=B, training_rounds=R)
    buf = _rollout_buffer(states, action, reward, terminated, truncated, A)
    pl = pearl_b200.B200ProximalPolicyOptimization(state_dim=obs, n_actions=A, actor_hidden_dims=[256, 256], critic_hidden_dims=[256, 256],
                                                   training_rounds=R, batch_size=B, actor_learning_rate=3e-4, critic_learning_rate=3e-4,
                                                   discount_factor=0.99, epsilon=0.2, trace_decay_param=0.95, entropy_bonus_scaling=0.01)
    pl.load_parameters(flat(orc.actor), flat(orc.critic))
    random.seed(5)
    trace = {}
    rep = pl.learn(buf, trace=trace)
    random.seed(5)
    t = torch.from_numpy
    otrace = {}
    orep, opre = orc.learn(t(states[:n]), t(action), t(reward), t(terminated), t(truncated), t(states[n]), trace=otrace)
    assert trace["idx"].tolist() == otrace["idx"]
    pre = pl.last_preprocess
    _close(pre["values"].cpu().numpy(), opre["values"].numpy(), "values", rtol=2e-5)
    _close(pre["action_probs"].cpu().numpy(), opre["action_probs"].numpy(), "action_probs", rtol=2e-5)
    _close(pre["gae"].cpu().numpy(), opre["gae"].numpy(), "gae", rtol=5e-5)
    _close(rep["actor_loss"], orep["actor_loss"], "actor_loss")
    _close(rep["critic_loss"], orep["critic_loss"], "critic_loss")
    _close_params(pl.actor_params.cpu().numpy(), flat(orc.actor).numpy(), "actor", 3e-4, R)
    _close_params(pl.critic_params.cpu().numpy(), flat(orc.critic).numpy(), "critic", 3e-4, R)


def test_ppo_rejects_bad_inputs():
    import pearl_b200
    pl = pearl_b200.B200ProximalPolicyOptimization(state_dim=4, n_actions=3, actor_hidden_dims=[8, 8], critic_hidden_dims=[8, 8])
    assert pl.learn(pearl_b200.B200ReplayBuffer(8)) == {}
    with pytest.raises(ValueError):
        pearl_b200.B200ProximalPolicyOptimization(state_dim=4, actor_hidden_dims=[8, 8], critic_hidden_dims=[8, 8])
    with pytest.raises(NotImplementedError):
        pearl_b200.B200ProximalPolicyOptimization(state_dim=4, n_actions=3, actor_hidden_dims=[8], critic_hidden_dims=[8, 8])
