"""TD3 and DDPG on the GPU against the recordings of the reference (tests/golden/{td3,ddpg}_small.npz: same indices from
CPython's `random.sample`, same `torch.normal` target noise) and against oracle/td3_oracle.py on a larger shape.
Tolerance: elementwise 1e-4 (tests/_tol.py)."""
import os
import random

import numpy as np
import pytest
import torch

from _tol import close as _close, close_params as _close_params
from conftest import GOLDEN
from oracle.pearl_oracle import flat
from oracle.td3_oracle import OracleTD3

pytestmark = pytest.mark.gpu


def _fill(buf, st, ac, rw, ns, term):
    n = st.shape[0]
    buf.push_batch(torch.from_numpy(st), torch.from_numpy(ac), torch.from_numpy(rw), torch.from_numpy(ns),
                   torch.from_numpy(term), torch.zeros(n, dtype=torch.bool))


@pytest.mark.parametrize("kind", ["td3", "ddpg"])
def test_td3_matches_reference_recording(kind):
    import pearl_b200
    fx = np.load(os.path.join(GOLDEN, f"{kind}_small.npz"))
    R, B = int(fx["rounds"]), int(fx["batch"])
    buf = pearl_b200.B200ReplayBuffer(int(fx["n"]))
    buf.is_action_continuous = True
    _fill(buf, fx["state"], fx["action"], fx["reward"], fx["next_state"], fx["terminated"])
    cls = pearl_b200.B200TD3 if kind == "td3" else pearl_b200.B200DeepDeterministicPolicyGradient
    pl = cls(state_dim=int(fx["obs"]), low=fx["low"], high=fx["high"], actor_hidden_dims=[32, 32], critic_hidden_dims=[32, 32],
             training_rounds=R, batch_size=B, actor_learning_rate=float(fx["actor_lr"]), critic_learning_rate=float(fx["critic_lr"]),
             actor_soft_update_tau=float(fx["actor_tau"]), critic_soft_update_tau=float(fx["critic_tau"]), discount_factor=float(fx["gamma"]))
    pl.load_parameters(fx["init_actor"], fx["init_q1"], fx["init_q2"], fx["init_actor_t"], fx["init_q1t"], fx["init_q2t"])
    random.seed(61 if kind == "td3" else 62)            # the state the recording sampled from
    trace = {}
    rep = pl.learn(buf, noise=torch.from_numpy(fx["noise"]) if kind == "td3" else None, trace=trace)
    assert np.array_equal(trace["idx"].numpy(), fx["idx"])
    _close(rep["actor_loss"], fx["actor_loss"], "actor_loss")
    _close(rep["critic_loss"], fx["critic_loss"], "critic_loss")
    pc = pl.critic_params.numel() // 2
    _close(pl.actor_params.cpu().numpy(), fx["actor_after"], "actor")
    _close(pl.actor_target_params.cpu().numpy(), fx["actor_t_after"], "actor target")
    _close(pl.critic_params[:pc].cpu().numpy(), fx["q1_after"], "q1")
    _close(pl.critic_params[pc:].cpu().numpy(), fx["q2_after"], "q2")
    _close(pl.critic_target_params[:pc].cpu().numpy(), fx["q1t_after"], "q1 target")
    _close(pl.critic_target_params[pc:].cpu().numpy(), fx["q2t_after"], "q2 target")


@pytest.mark.parametrize("graph", [True, False])
def test_td3_larger_shape_against_oracle_two_learn_calls(graph):
    """obs 17, 6 actions, [256, 256] networks, batch 256, delayed updates with freq 3 across two learn() calls (the actor's
    own AdamW step count and the update phase carry over), CUDA-graph replay and plain launches."""
    import pearl_b200
    obs, act, n, B, R = 17, 6, 3000, 256, 4
    rng = np.random.Generator(np.random.PCG64(33))
    st = rng.standard_normal((n, obs)).astype(np.float32); ns = rng.standard_normal((n, obs)).astype(np.float32)
    rw = rng.standard_normal(n).astype(np.float32); term = rng.random(n) < 0.03
    low, high = -np.ones(act, dtype=np.float32), np.ones(act, dtype=np.float32) * 2
    ac = rng.uniform(low, high, size=(n, act)).astype(np.float32)
    buf = pearl_b200.B200ReplayBuffer(n)
    buf.is_action_continuous = True
    _fill(buf, st, ac, rw, ns, term)
    pl = pearl_b200.B200TD3(state_dim=obs, low=low, high=high, actor_hidden_dims=[256, 256], critic_hidden_dims=[256, 256],
                            training_rounds=R, batch_size=B, actor_learning_rate=3e-4, critic_learning_rate=3e-4, actor_update_freq=3,
                            actor_update_noise=0.2, actor_update_noise_clip=0.5, seed=4)
    pl.use_cuda_graph = graph
    pc = pl.critic_params.numel() // 2
    init = dict(actor=pl.actor_params.cpu().numpy(), actor_t=pl.actor_target_params.cpu().numpy(), q1=pl.critic_params[:pc].cpu().numpy(),
                q2=pl.critic_params[pc:].cpu().numpy(), q1t=pl.critic_target_params[:pc].cpu().numpy(), q2t=pl.critic_target_params[pc:].cpu().numpy())
    orc = OracleTD3(obs, act, (256, 256), (256, 256), low, high, actor_lr=3e-4, critic_lr=3e-4, actor_update_freq=3, init=init)
    g = torch.Generator().manual_seed(8)
    t = torch.from_numpy
    random.seed(77)
    al, cl, gl, gc = [], [], [], []
    for call in range(2):
        noise = torch.randn((R, B, act), generator=g) * 0.2
        trace = {}
        state = random.getstate()
        rep = pl.learn(buf, noise=noise, trace=trace)
        gl += rep["actor_loss"]; gc += rep["critic_loss"]
        random.setstate(state)
        for r in range(R):
            idx = random.sample(range(n), B)
            assert idx == trace["idx"][r].tolist()
            b = dict(state=t(st[idx]), action=t(ac[idx]), reward=t(rw[idx]), next_state=t(ns[idx]), terminated=t(term[idx]))
            orc.training_steps += 1
            out = orc.learn_batch(b, noise[r])
            al.append(out["actor_loss"]); cl.append(out["critic_loss"])
    _close(gl, al, "actor_loss")
    _close(gc, cl, "critic_loss")
    _close_params(pl.actor_params.cpu().numpy(), flat(orc.actor).numpy(), "actor", 3e-4, 2 * R)
    _close_params(pl.actor_target_params.cpu().numpy(), flat(orc.actor_t).numpy(), "actor target", 3e-4, 2 * R)
    _close_params(pl.critic_params[:pc].cpu().numpy(), flat(orc.q[0]).numpy(), "q1", 3e-4, 2 * R)
    _close_params(pl.critic_params[pc:].cpu().numpy(), flat(orc.q[1]).numpy(), "q2", 3e-4, 2 * R)
    _close_params(pl.critic_target_params[:pc].cpu().numpy(), flat(orc.qt[0]).numpy(), "q1 target", 3e-4, 2 * R)
