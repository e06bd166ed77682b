"""The PPO restatement (oracle/ppo_oracle.py: OraclePPO) against the recording of the reference's
PearlAgent(ProximalPolicyOptimization, PPOReplayBuffer).learn() (tests/golden/ppo_small.npz)."""
import os
import random

import numpy as np
import torch

from conftest import GOLDEN
from oracle.pearl_oracle import flat
from oracle.ppo_oracle import OraclePPO


def make_oracle(fx):
    return OraclePPO(int(fx["obs"]), int(fx["n_act"]), (32, 32), (32, 32), actor_lr=float(fx["actor_lr"]),
                     critic_lr=float(fx["critic_lr"]), gamma=float(fx["gamma"]), epsilon=float(fx["epsilon"]),
                     trace_decay=float(fx["lam"]), entropy_bonus=float(fx["beta"]), batch_size=int(fx["batch"]),
                     training_rounds=int(fx["rounds"]), init_actor=fx["init_actor"], init_critic=fx["init_critic"])


def test_ppo_oracle_reproduces_reference():
    torch.set_num_threads(1)
    fx = np.load(os.path.join(GOLDEN, "ppo_small.npz"))
    n = int(fx["n"])
    orc = make_oracle(fx)
    t = torch.from_numpy
    random.seed(53)
    trace = {}
    rep, pre = orc.learn(t(fx["states"][:n]), t(fx["action"]), t(fx["reward"]), t(fx["terminated"]), t(fx["truncated"]),
                         t(fx["states"][n]), trace=trace)
    assert trace["idx"] == fx["idx"].tolist()
    tol = dict(rtol=5e-6, atol=5e-7)
    np.testing.assert_allclose(pre["gae"].numpy(), fx["gae"], **tol)
    np.testing.assert_allclose(pre["lam_return"].numpy(), fx["lam_return"], **tol)
    np.testing.assert_allclose(pre["action_probs"].numpy(), fx["action_probs"], **tol)
    np.testing.assert_allclose(rep["actor_loss"], fx["actor_loss"], rtol=5e-6)
    np.testing.assert_allclose(rep["critic_loss"], fx["critic_loss"], rtol=5e-6)
    np.testing.assert_allclose(flat(orc.actor).numpy(), fx["actor_after"], **tol)
    np.testing.assert_allclose(flat(orc.critic).numpy(), fx["critic_after"], **tol)
