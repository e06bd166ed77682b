"""The SAC restatement (oracle/sac_oracle.py) against the recording of the reference's
PearlAgent(ContinuousSoftActorCritic, BasicReplayBuffer).learn() (tests/golden/sac_small.npz)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle.pearl_oracle import flat
from oracle.sac_oracle import OracleSAC


def make_oracle(fx):
    init = {k: fx[f"init_{k}"] for k in ("actor", "q1", "q2", "q1t", "q2t")}
    return OracleSAC(int(fx["obs"]), int(fx["act"]), (32, 32), (32, 32), fx["low"], fx["high"], actor_lr=float(fx["actor_lr"]),
                     critic_lr=float(fx["critic_lr"]), gamma=float(fx["gamma"]), tau=float(fx["tau"]), init=init)


def batch_of(fx, idx):
    t = lambda k: torch.from_numpy(fx[k][idx])
    return dict(state=t("state"), action=t("action"), reward=t("reward"), next_state=t("next_state"), terminated=t("terminated"))


def test_sac_oracle_reproduces_reference():
    torch.set_num_threads(1)
    fx = np.load(os.path.join(GOLDEN, "sac_small.npz"))
    orc = make_oracle(fx)
    for r in range(int(fx["rounds"])):
        out = orc.learn_batch(batch_of(fx, fx["idx"][r]), torch.from_numpy(fx["noise"][2 * r]), torch.from_numpy(fx["noise"][2 * r + 1]))
        np.testing.assert_allclose(out["actor_loss"], fx["actor_loss"][r], rtol=5e-6)
        np.testing.assert_allclose(out["critic_loss"], fx["critic_loss"][r], rtol=5e-6)
        np.testing.assert_allclose(out["entropy_coef"], fx["entropy_loss"][r], rtol=5e-6, atol=1e-7)
    tol = dict(rtol=5e-6, atol=5e-7)
    np.testing.assert_allclose(flat(orc.actor).numpy(), fx["actor_after"], **tol)
    np.testing.assert_allclose(flat(orc.q[0]).numpy(), fx["q1_after"], **tol)
    np.testing.assert_allclose(flat(orc.q[1]).numpy(), fx["q2_after"], **tol)
    np.testing.assert_allclose(flat(orc.qt[0]).numpy(), fx["q1t_after"], **tol)
    np.testing.assert_allclose(flat(orc.qt[1]).numpy(), fx["q2t_after"], **tol)
    np.testing.assert_allclose(orc.log_alpha.detach().numpy(), fx["log_alpha_after"], **tol)
