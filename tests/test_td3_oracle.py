"""The CPU restatement of TD3 / DDPG (oracle/td3_oracle.py) against the recordings of the reference's own
PearlAgent(TD3 | DeepDeterministicPolicyGradient).learn() (tests/golden/{td3,ddpg}_small.npz, oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.pearl_oracle import flat
from oracle.td3_oracle import OracleTD3


@pytest.mark.parametrize("kind", ["td3", "ddpg"])
def test_oracle_td3_reproduces_the_reference_recording(kind):
    fx = np.load(os.path.join(GOLDEN, f"{kind}_small.npz"))
    orc = OracleTD3(int(fx["obs"]), int(fx["act"]), (32, 32), (32, 32), fx["low"], fx["high"], actor_lr=float(fx["actor_lr"]),
                    critic_lr=float(fx["critic_lr"]), gamma=float(fx["gamma"]), actor_tau=float(fx["actor_tau"]),
                    critic_tau=float(fx["critic_tau"]), actor_update_freq=int(fx["freq"]), noise_clip=float(fx["noise_clip"]),
                    init={k: fx[f"init_{k}"] for k in ("actor", "actor_t", "q1", "q2", "q1t", "q2t")})
    t = torch.from_numpy
    al, cl = [], []
    for r in range(int(fx["rounds"])):
        ix = fx["idx"][r].astype(np.int64)
        b = dict(state=t(fx["state"][ix]), action=t(fx["action"][ix]), reward=t(fx["reward"][ix]), next_state=t(fx["next_state"][ix]),
                 terminated=t(fx["terminated"][ix]))
        orc.training_steps += 1
        out = orc.learn_batch(b, t(fx["noise"][r]) if kind == "td3" else None)
        al.append(out["actor_loss"]); cl.append(out["critic_loss"])
    np.testing.assert_allclose(al, fx["actor_loss"], rtol=5e-6, atol=1e-7)
    np.testing.assert_allclose(cl, fx["critic_loss"], rtol=5e-6, atol=1e-7)
    for name, net in (("actor", orc.actor), ("actor_t", orc.actor_t), ("q1", orc.q[0]), ("q2", orc.q[1]), ("q1t", orc.qt[0]), ("q2t", orc.qt[1])):
        np.testing.assert_allclose(flat(net).numpy(), fx[f"{name}_after"], rtol=5e-6, atol=1e-7, err_msg=name)
