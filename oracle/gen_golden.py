#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

TEST INFRASTRUCTURE.  This script is the only place in the repo that imports
`/root/reference/pearl` (read-only, via the test-only stubs in oracle/stubs/).
It cannot run on the GPU box (no /root/reference there); its outputs are
committed as small .npz/.json fixtures and everything else (oracle tests,
GPU parity tests, smoke) reads those.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

What is recorded, per case (see `run_dqn_case`):
  * the pushed transitions (quantised to a 1/256 grid, stored as int16, so the
    fixture is small and exactly representable in fp32),
  * the CPython `random` state before and after `learn()`
    (reference call site: pearl/replay_buffers/tensor_based_replay_buffer.py:276),
  * per training round: the sampled logical indices (0 = oldest element of the
    deque), Q(s,a) (`state_action_values`, deep_td_learning.py:342), the
    Bellman target y (`expected_state_action_values`, :313-317) and the
    reported "loss" (mean |q - y|, :358-360),
  * parameter snapshots of `_Q` / `_Q_target` after selected rounds, and the
    final AdamW(amsgrad) state (torch/optim/adam.py `_single_tensor_adam`).
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(HERE, "stubs"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle.synth import make_transitions  # noqa: E402

from pearl.action_representation_modules.one_hot_action_representation_module import (  # noqa: E402
    OneHotActionTensorRepresentationModule,
)
from pearl.pearl_agent import PearlAgent  # noqa: E402
from pearl.policy_learners.sequential_decision_making.deep_q_learning import (  # noqa: E402
    DeepQLearning,
)
from pearl.policy_learners.sequential_decision_making.double_dqn import DoubleDQN  # noqa: E402
from pearl.replay_buffers.basic_replay_buffer import BasicReplayBuffer  # noqa: E402
from pearl.utils.instantiations.spaces.discrete_action import DiscreteActionSpace  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def state_words(st) -> np.ndarray:
    """random.getstate() -> uint32[625] (624 MT words + position index)."""
    assert st[0] == 3 and st[2] is None
    return np.asarray(st[1], dtype=np.uint64).astype(np.uint32)


def flat_params(module) -> np.ndarray:
    return np.concatenate([p.detach().numpy().ravel() for p in module.parameters()])


def run_dqn_case(name, *, obs, n_act, hidden, capacity, n_push, batch, rounds,
                 target_update_freq, tau, double, seed, data_seed, dynamic=False,
                 snap_rounds=(1, 2, 5), lr=1e-3, gamma=0.99, learn_calls=1):
    torch.manual_seed(seed)
    random.seed(seed)
    torch.set_num_threads(1)
    cls = DoubleDQN if double else DeepQLearning
    full_space = DiscreteActionSpace(actions=list(torch.arange(n_act).view(-1, 1)))
    learner = cls(
        state_dim=obs,
        action_space=full_space,
        hidden_dims=list(hidden),
        learning_rate=lr,
        discount_factor=gamma,
        training_rounds=rounds,
        batch_size=batch,
        target_update_freq=target_update_freq,
        soft_update_tau=tau,
        action_representation_module=OneHotActionTensorRepresentationModule(n_act),
    )
    buf = BasicReplayBuffer(capacity)
    agent = PearlAgent(policy_learner=learner, replay_buffer=buf, device_id=-1)
    assert str(agent.device) == "cpu"

    data = make_transitions(n_push, obs, n_act, seed=data_seed, dynamic=dynamic)
    for i in range(n_push):
        if dynamic:
            ids = data["next_avail_ids"][i, : data["next_avail_n"][i]]
            nxt = DiscreteActionSpace(actions=[torch.tensor([int(a)]) for a in ids])
        else:
            nxt = full_space
        buf.push(
            state=torch.from_numpy(data["state"][i]),
            action=torch.tensor(int(data["action"][i])),
            reward=float(data["reward"][i]),
            terminated=bool(data["terminated"][i]),
            truncated=bool(data["truncated"][i]),
            curr_available_actions=full_space,
            next_state=torch.from_numpy(data["next_state"][i]),
            next_available_actions=nxt,
            max_number_actions=n_act,
        )

    init_q = flat_params(learner._Q).copy()
    init_qt = flat_params(learner._Q_target).copy()

    # --- instrument the reference (observation only; arithmetic untouched) ----
    rec = {"idx": [], "q": [], "y": [], "mae": []}
    snaps_q, snaps_qt = {}, {}
    orig_sample = buf.sample
    orig_loss = learner.loss
    orig_learn_batch = learner.learn_batch

    def sample_spy(batch_size):
        pos = {id(t): j for j, t in enumerate(buf.memory)}
        # same call the reference makes, with the transitions identified afterwards
        st = random.getstate()
        picked = random.sample(buf.memory, batch_size)
        rec["idx"].append([pos[id(t)] for t in picked])
        random.setstate(st)
        return orig_sample(batch_size)

    def loss_spy(batch_, predictions):
        loss, y = orig_loss(batch_, predictions)
        rec["q"].append(predictions.detach().numpy().copy())
        rec["y"].append(y.detach().numpy().copy())
        return loss, y

    def learn_batch_spy(batch_):
        out = orig_learn_batch(batch_)
        rec["mae"].append(out["loss"])
        r = len(rec["mae"])
        if r in snap_rounds or r == rounds * learn_calls:
            snaps_q[r] = flat_params(learner._Q).copy()
            snaps_qt[r] = flat_params(learner._Q_target).copy()
        return out

    buf.sample = sample_spy
    learner.loss = loss_spy
    learner.learn_batch = learn_batch_spy

    state0 = state_words(random.getstate())
    reports = []
    for _ in range(learn_calls):
        reports.append(agent.learn())
    state1 = state_words(random.getstate())
    total = rounds * learn_calls
    assert len(rec["mae"]) == total
    assert [v for r in reports for v in r["loss"]] == rec["mae"]

    opt_state = learner.optimizer.state_dict()["state"]
    exp_avg = np.concatenate([opt_state[i]["exp_avg"].numpy().ravel() for i in range(len(opt_state))])
    exp_avg_sq = np.concatenate([opt_state[i]["exp_avg_sq"].numpy().ravel() for i in range(len(opt_state))])
    max_sq = np.concatenate([opt_state[i]["max_exp_avg_sq"].numpy().ravel() for i in range(len(opt_state))])
    step = float(opt_state[0]["step"])
    assert step == total

    cfg = dict(name=name, obs=obs, n_act=n_act, hidden=list(hidden), capacity=capacity,
               n_push=n_push, batch=batch, rounds=rounds, learn_calls=learn_calls,
               target_update_freq=target_update_freq, tau=tau, double=bool(double),
               seed=seed, data_seed=data_seed, dynamic=bool(dynamic), lr=lr, gamma=gamma,
               betas=[0.9, 0.999], eps=1e-8, weight_decay=0.01, amsgrad=True,
               snap_rounds=sorted(snaps_q.keys()), torch=torch.__version__,
               reference="facebookresearch/Pearl @ 48f1fbb")
    out = dict(
        config=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8),
        state_q=data["state_q"], next_state_q=data["next_state_q"],
        action=data["action"].astype(np.int16), reward_q=data["reward_q"],
        terminated=data["terminated"].astype(np.uint8),
        truncated=data["truncated"].astype(np.uint8),
        mt_state_before=state0, mt_state_after=state1,
        idx=np.asarray(rec["idx"], dtype=np.int32),
        q=np.asarray(rec["q"], dtype=np.float32),
        y=np.asarray(rec["y"], dtype=np.float32),
        mae=np.asarray(rec["mae"], dtype=np.float64),
        init_q=init_q, init_q_target=init_qt,
        exp_avg=exp_avg, exp_avg_sq=exp_avg_sq, max_exp_avg_sq=max_sq,
    )
    if dynamic:
        out["next_avail_ids"] = data["next_avail_ids"].astype(np.int16)
        out["next_avail_n"] = data["next_avail_n"].astype(np.int16)
    for r in snaps_q:
        out[f"q_after_{r}"] = snaps_q[r]
        out[f"qt_after_{r}"] = snaps_qt[r]
    path = os.path.join(GOLDEN, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB  mae[0]={rec['mae'][0]:.6f} "
          f"mae[-1]={rec['mae'][-1]:.6f}")


def gen_random_sample_vectors():
    """Known-answer vectors for CPython's MT19937 + random.sample
    (cpython Lib/random.py:242-250,359-452; Modules/_randommodule.c)."""
    out = []
    for seed, n, k, reps in [
        (1234, 10**6, 256, 3),     # set branch, 20-bit draws (cfg2)
        (1234, 1045, 256, 2),      # pool branch, largest n for k=256
        (1234, 1046, 256, 2),      # set branch, smallest n for k=256
        (7, 4_000_000, 256, 2),    # cfg5 size
        (99, 10_000, 32, 4),       # cfg1
        (99, 277, 32, 2), (99, 278, 32, 2),
        (5, 256, 256, 2),          # k == n : a permutation
        (5, 1, 1, 3),              # n == 1: getrandbits(1) rejection loop
        (5, 65_536, 256, 2),       # n a power of two
        (2**40 + 12345, 123_457, 512, 2),  # big-int seed (init_by_array, 2 keys), k=512
        (0, 300, 200, 2),          # pool branch with k close to n
        (31337, 5000, 4000, 1),    # large k: set branch, many duplicate retries
    ]:
        random.seed(seed)
        st0 = state_words(random.getstate())
        first_words = [random.getrandbits(32) for _ in range(4)]
        random.seed(seed)
        samples = [random.sample(range(n), k) for _ in range(reps)]
        st1 = state_words(random.getstate())
        tail = random.getrandbits(32)
        out.append(dict(seed=seed, n=n, k=k, reps=reps, first_words=first_words,
                        samples=samples, next_word_after=tail,
                        state_before_xor=int(np.bitwise_xor.reduce(st0[:624])),
                        state_after_index=int(st1[624]),
                        state_after_xor=int(np.bitwise_xor.reduce(st1[:624]))))
    with open(os.path.join(GOLDEN, "random_sample_kat.json"), "w") as f:
        json.dump(dict(source="CPython %s random.seed/getrandbits/sample" % sys.version.split()[0],
                       cases=out), f)
    print("random_sample_kat.json:", len(out), "cases")


def gen_ppo_gae_golden():
    """GAE / lambda-return vectors from the reference's own preprocess_replay_buffer (ppo.py:201-293)."""
    from pearl.policy_learners.sequential_decision_making.ppo import PPOReplayBuffer, ProximalPolicyOptimization
    torch.manual_seed(31)
    random.seed(31)
    obs, n_act, n = 6, 4, 700
    space = DiscreteActionSpace(actions=list(torch.arange(n_act).view(-1, 1)))
    pl = ProximalPolicyOptimization(state_dim=obs, action_space=space, actor_hidden_dims=[16, 16], critic_hidden_dims=[16, 16],
                                    training_rounds=1, batch_size=32, epsilon=0.1, discount_factor=0.97, trace_decay_param=0.9,
                                    action_representation_module=OneHotActionTensorRepresentationModule(n_act))
    buf = PPOReplayBuffer(n)
    agent = PearlAgent(policy_learner=pl, replay_buffer=buf, device_id=-1)
    rng = np.random.Generator(np.random.PCG64(5))
    states = (np.rint(rng.standard_normal((n + 1, obs)) * 256) / 256).astype(np.float32)
    rewards = (np.rint(rng.standard_normal(n) * 256) / 256).astype(np.float32)
    terminated = np.zeros(n, dtype=bool); truncated = np.zeros(n, dtype=bool)
    terminated[49::50] = True
    truncated[[120, 333, 512]] = True
    terminated[[7, 8]] = True          # adjacent episode ends
    for i in range(n):
        buf.push(state=torch.from_numpy(states[i]), action=torch.tensor(i % n_act), reward=float(rewards[i]),
                 terminated=bool(terminated[i]), truncated=bool(truncated[i]), curr_available_actions=space,
                 next_state=torch.from_numpy(states[i + 1]), next_available_actions=space, max_number_actions=n_act)
    with torch.no_grad():
        values = pl._critic(torch.from_numpy(states[:n])).reshape(n).numpy().copy()
        last_next_value = float(pl._critic(torch.from_numpy(states[n:n + 1]))[0])
    pl.preprocess_replay_buffer(buf)
    gae = np.asarray([float(t.gae) for t in buf.memory], dtype=np.float32)
    lam = np.asarray([float(t.lam_return) for t in buf.memory], dtype=np.float32)
    # the closed form of the reference's unit test (test_ppo.py:48-115): gamma 0.6, lambda 0.5, rewards 4,6,5
    v = np.asarray([0.37, -1.25, 0.5, 2.0], dtype=np.float32)
    g2 = np.float32(5) + np.float32(0.6) * v[3] - v[2]
    g1 = np.float32(6) + np.float32(0.6) * v[2] - v[1] + np.float32(0.6 * 0.5) * g2
    g0 = np.float32(4) + np.float32(0.6) * v[1] - v[0] + np.float32(0.6 * 0.5) * g1
    np.savez_compressed(os.path.join(GOLDEN, "ppo_gae.npz"), values=values, last_next_value=np.float32(last_next_value),
                        reward=rewards, terminated=terminated, truncated=truncated, gae=gae, lam_return=lam,
                        gamma=0.97, lam=0.9, kat_v=v, kat_gae=np.asarray([g0, g1, g2], dtype=np.float32),
                        kat_reward=np.asarray([4, 6, 5], dtype=np.float32))
    print("ppo_gae.npz:", n, "transitions; gae[0..2] =", gae[:3])


def gen_ppo_learn_golden():
    """PPO end to end: PearlAgent(ProximalPolicyOptimization, PPOReplayBuffer).learn() (preprocess + clipped-surrogate
    actor steps + critic steps) with the sampled indices recorded."""
    from pearl.policy_learners.sequential_decision_making.ppo import PPOReplayBuffer, ProximalPolicyOptimization
    torch.manual_seed(53)
    random.seed(53)
    torch.set_num_threads(1)
    obs, n_act, n, B, rounds = 10, 5, 400, 48, 6
    space = DiscreteActionSpace(actions=list(torch.arange(n_act).view(-1, 1)))
    hp = dict(actor_lr=3e-4, critic_lr=1e-3, epsilon=0.2, gamma=0.97, lam=0.9, beta=0.02)
    pl = ProximalPolicyOptimization(state_dim=obs, action_space=space, actor_hidden_dims=[32, 32], critic_hidden_dims=[32, 32],
                                    training_rounds=rounds, batch_size=B, epsilon=hp["epsilon"], discount_factor=hp["gamma"],
                                    trace_decay_param=hp["lam"], entropy_bonus_scaling=hp["beta"],
                                    actor_learning_rate=hp["actor_lr"], critic_learning_rate=hp["critic_lr"],
                                    action_representation_module=OneHotActionTensorRepresentationModule(n_act))
    buf = PPOReplayBuffer(n)
    agent = PearlAgent(policy_learner=pl, replay_buffer=buf, device_id=-1)
    rng = np.random.Generator(np.random.PCG64(15))
    q8 = lambda x: (np.rint(x * 256) / 256).astype(np.float32)
    states = q8(rng.standard_normal((n + 1, obs)))
    rewards = q8(rng.standard_normal(n))
    actions = rng.integers(0, n_act, size=n).astype(np.int64)
    terminated = np.zeros(n, dtype=bool); truncated = np.zeros(n, dtype=bool)
    terminated[39::40] = True
    truncated[[100, 333]] = True
    for i in range(n):
        buf.push(state=torch.from_numpy(states[i]), action=torch.tensor(int(actions[i])), reward=float(rewards[i]),
                 terminated=bool(terminated[i]), truncated=bool(truncated[i]), curr_available_actions=space,
                 next_state=torch.from_numpy(states[i + 1]), next_available_actions=space, max_number_actions=n_act)
    fl = lambda m: np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
    init_actor, init_critic = fl(pl._actor), fl(pl._critic)
    idxs, pre = [], {}
    orig_sample = buf.sample

    def sample_spy(k):
        if not pre:     # the agent clears an on-policy buffer after learning: record the preprocessing results now
            pre["gae"] = np.asarray([float(t.gae) for t in buf.memory], dtype=np.float32)
            pre["lam"] = np.asarray([float(t.lam_return) for t in buf.memory], dtype=np.float32)
            pre["apo"] = np.asarray([float(t.action_probs) for t in buf.memory], dtype=np.float32)
        pos = {id(t): j for j, t in enumerate(buf.memory)}
        stt = random.getstate()
        idxs.append([pos[id(t)] for t in random.sample(buf.memory, k)])
        random.setstate(stt)
        return orig_sample(k)
    buf.sample = sample_spy
    rep = agent.learn()
    gae, lam, apo = pre["gae"], pre["lam"], pre["apo"]
    assert gae.shape == (n,)
    np.savez_compressed(os.path.join(GOLDEN, "ppo_small.npz"), obs=obs, n_act=n_act, n=n, batch=B, rounds=rounds, **hp,
                        states=states, action=actions, reward=rewards, terminated=terminated, truncated=truncated,
                        idx=np.asarray(idxs, dtype=np.int32), gae=gae, lam_return=lam, action_probs=apo,
                        actor_loss=np.asarray(rep["actor_loss"]), critic_loss=np.asarray(rep["critic_loss"]),
                        init_actor=init_actor, init_critic=init_critic, actor_after=fl(pl._actor), critic_after=fl(pl._critic))
    print("ppo_small.npz: actor_loss", rep["actor_loss"][:2], "critic_loss", rep["critic_loss"][:2])


def gen_sac_golden():
    """Continuous SAC: PearlAgent(ContinuousSoftActorCritic, BasicReplayBuffer).learn() with the two
    Normal.rsample noise draws per step recorded (torch.distributions.normal._standard_normal)."""
    import torch.distributions.normal as tdn
    from pearl.policy_learners.sequential_decision_making.soft_actor_critic_continuous import ContinuousSoftActorCritic
    from pearl.utils.instantiations.spaces.box_action import BoxActionSpace
    torch.manual_seed(41)
    random.seed(41)
    torch.set_num_threads(1)
    obs, act, n, B, rounds = 11, 3, 300, 64, 6
    low, high = torch.tensor([-0.4, -0.8, -0.4]), torch.tensor([0.4, 0.8, 1.2])
    space = BoxActionSpace(low=low, high=high)
    pl = ContinuousSoftActorCritic(state_dim=obs, action_space=space, actor_hidden_dims=[32, 32], critic_hidden_dims=[32, 32],
                                   training_rounds=rounds, batch_size=B, actor_learning_rate=3e-4, critic_learning_rate=5e-4,
                                   critic_soft_update_tau=0.05, discount_factor=0.98, entropy_autotune=True)
    buf = BasicReplayBuffer(n)
    agent = PearlAgent(policy_learner=pl, replay_buffer=buf, device_id=-1)
    rng = np.random.Generator(np.random.PCG64(9))
    q8 = lambda x: (np.rint(x * 256) / 256).astype(np.float32)
    st, ns, rw = q8(rng.standard_normal((n, obs))), q8(rng.standard_normal((n, obs))), q8(rng.standard_normal(n))
    ac = q8(rng.uniform(low.numpy(), high.numpy(), size=(n, act)))
    term = rng.random(n) < 0.05
    for i in range(n):
        buf.push(state=torch.from_numpy(st[i]), action=torch.from_numpy(ac[i]), reward=float(rw[i]), terminated=bool(term[i]),
                 truncated=False, curr_available_actions=space, next_state=torch.from_numpy(ns[i]), next_available_actions=space)
    fl = lambda m: np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
    init = dict(actor=fl(pl._actor), q1=fl(pl._critic._critic_1), q2=fl(pl._critic._critic_2),
                q1t=fl(pl._critic_target._critic_1), q2t=fl(pl._critic_target._critic_2))
    noises, idxs = [], []
    orig_sn = tdn._standard_normal

    def sn_spy(shape, dtype, device):
        x = orig_sn(shape, dtype, device)
        noises.append(x.numpy().copy())
        return x
    tdn._standard_normal = sn_spy
    orig_sample = buf.sample

    def sample_spy(k):
        pos = {id(t): j for j, t in enumerate(buf.memory)}
        stt = random.getstate()
        idxs.append([pos[id(t)] for t in random.sample(buf.memory, k)])
        random.setstate(stt)
        return orig_sample(k)
    buf.sample = sample_spy
    rep = agent.learn()
    tdn._standard_normal = orig_sn
    assert len(noises) == 2 * rounds
    out = dict(obs=obs, act=act, n=n, batch=B, rounds=rounds, low=low.numpy(), high=high.numpy(), actor_lr=3e-4, critic_lr=5e-4,
               tau=0.05, gamma=0.98, state=st, next_state=ns, reward=rw, action=ac, terminated=term,
               idx=np.asarray(idxs, dtype=np.int32), noise=np.asarray(noises, dtype=np.float32),
               actor_loss=np.asarray(rep["actor_loss"]), critic_loss=np.asarray(rep["critic_loss"]),
               entropy_loss=np.asarray([float(x) for x in rep["entropy_coef"]]),
               actor_after=fl(pl._actor), q1_after=fl(pl._critic._critic_1), q2_after=fl(pl._critic._critic_2),
               q1t_after=fl(pl._critic_target._critic_1), q2t_after=fl(pl._critic_target._critic_2),
               log_alpha_after=pl._log_entropy.detach().numpy().copy(), **{f"init_{k}": v for k, v in init.items()})
    np.savez_compressed(os.path.join(GOLDEN, "sac_small.npz"), **out)
    print("sac_small.npz: actor_loss", rep["actor_loss"][:2], "critic_loss", rep["critic_loss"][:2])


def gen_td3_golden(kind: str):
    """TD3 / DDPG: PearlAgent(TD3 | DeepDeterministicPolicyGradient, BasicReplayBuffer).learn() with the sampled indices and
    (TD3) the `torch.normal` target-noise draws recorded."""
    from pearl.policy_learners.exploration_modules.common.no_exploration import NoExploration
    from pearl.policy_learners.sequential_decision_making.ddpg import DeepDeterministicPolicyGradient
    from pearl.policy_learners.sequential_decision_making.td3 import TD3
    from pearl.utils.instantiations.spaces.box_action import BoxActionSpace
    torch.manual_seed(61 if kind == "td3" else 62)
    random.seed(61 if kind == "td3" else 62)
    torch.set_num_threads(1)
    obs, act, n, B, rounds = 9, 3, 260, 48, 8
    low, high = torch.tensor([-0.5, -1.0, -0.25]), torch.tensor([0.5, 1.0, 1.25])
    space = BoxActionSpace(low=low, high=high)
    hp = dict(actor_lr=3e-4, critic_lr=6e-4, actor_tau=0.03, critic_tau=0.05, gamma=0.97)
    common = dict(state_dim=obs, action_space=space, actor_hidden_dims=[32, 32], critic_hidden_dims=[32, 32], training_rounds=rounds,
                  batch_size=B, actor_learning_rate=hp["actor_lr"], critic_learning_rate=hp["critic_lr"],
                  actor_soft_update_tau=hp["actor_tau"], critic_soft_update_tau=hp["critic_tau"], discount_factor=hp["gamma"],
                  exploration_module=NoExploration())
    if kind == "td3":
        hp.update(freq=2, noise_std=0.2, noise_clip=0.5)
        pl = TD3(actor_update_freq=2, actor_update_noise=0.2, actor_update_noise_clip=0.5, **common)
    else:
        hp.update(freq=1, noise_std=0.0, noise_clip=0.0)
        pl = DeepDeterministicPolicyGradient(**common)
    buf = BasicReplayBuffer(n)
    agent = PearlAgent(policy_learner=pl, replay_buffer=buf, device_id=-1)
    rng = np.random.Generator(np.random.PCG64(19))
    q8 = lambda x: (np.rint(x * 256) / 256).astype(np.float32)
    st, ns, rw = q8(rng.standard_normal((n, obs))), q8(rng.standard_normal((n, obs))), q8(rng.standard_normal(n))
    ac = q8(rng.uniform(low.numpy(), high.numpy(), size=(n, act)))
    term = rng.random(n) < 0.05
    for i in range(n):
        buf.push(state=torch.from_numpy(st[i]), action=torch.from_numpy(ac[i]), reward=float(rw[i]), terminated=bool(term[i]),
                 truncated=False, curr_available_actions=space, next_state=torch.from_numpy(ns[i]), next_available_actions=space)
    fl = lambda m: np.concatenate([p.detach().numpy().ravel() for p in m.parameters()])
    nets = lambda: dict(actor=fl(pl._actor), actor_t=fl(pl._actor_target), q1=fl(pl._critic._critic_1), q2=fl(pl._critic._critic_2),
                        q1t=fl(pl._critic_target._critic_1), q2t=fl(pl._critic_target._critic_2))
    init = nets()
    noises, idxs = [], []
    orig_normal = torch.normal

    def normal_spy(*a, **k):
        x = orig_normal(*a, **k)
        noises.append(x.numpy().copy())
        return x
    torch.normal = normal_spy
    orig_sample = buf.sample

    def sample_spy(k):
        pos = {id(t): j for j, t in enumerate(buf.memory)}
        stt = random.getstate()
        idxs.append([pos[id(t)] for t in random.sample(buf.memory, k)])
        random.setstate(stt)
        return orig_sample(k)
    buf.sample = sample_spy
    rep = agent.learn()
    torch.normal = orig_normal
    assert len(noises) == (rounds if kind == "td3" else 0)
    out = dict(kind=kind, obs=obs, act=act, n=n, batch=B, rounds=rounds, low=low.numpy(), high=high.numpy(), state=st, next_state=ns,
               reward=rw, action=ac, terminated=term, idx=np.asarray(idxs, dtype=np.int32),
               noise=np.asarray(noises, dtype=np.float32).reshape(len(noises), B, act) if noises else np.zeros((0, B, act), np.float32),
               actor_loss=np.asarray(rep["actor_loss"]), critic_loss=np.asarray(rep["critic_loss"]),
               **{f"init_{k}": v for k, v in init.items()}, **{f"{k}_after": v for k, v in nets().items()}, **hp)
    np.savez_compressed(os.path.join(GOLDEN, f"{kind}_small.npz"), **out)
    print(f"{kind}_small.npz: actor_loss", rep["actor_loss"][:3], "critic_loss", rep["critic_loss"][:2])


def gen_her_golden():
    """HindsightExperienceReplayBuffer: the deque contents (oldest first) after three episodes pushed through the reference."""
    from pearl.replay_buffers.sequential_decision_making.hindsight_experience_replay_buffer import HindsightExperienceReplayBuffer
    goal_dim, n_act, cap = 3, 4, 40
    space = DiscreteActionSpace(actions=list(torch.arange(n_act).view(-1, 1)))
    reward_fn = lambda s, a: float((s[:goal_dim] - s[-goal_dim:]).abs().sum() < 0.75) - 0.5      # noqa: E731
    terminated_fn = lambda s, a: bool((s[:goal_dim] - s[-goal_dim:]).abs().sum() < 0.25)        # noqa: E731
    rng = np.random.Generator(np.random.PCG64(77))
    q8 = lambda x: (np.rint(x * 8) / 8).astype(np.float32)
    lens = [7, 12, 5]
    total = sum(lens)
    st, ns = q8(rng.standard_normal((total, 2 * goal_dim))), q8(rng.standard_normal((total, 2 * goal_dim)))
    rw = q8(rng.standard_normal(total))
    ac = rng.integers(0, n_act, size=total).astype(np.int64)
    ends = np.zeros(total, dtype=bool); ends[np.cumsum(lens) - 1] = True
    trunc = np.zeros(total, dtype=bool); trunc[lens[0] - 1] = True          # the first episode is truncated, the others terminate
    term = ends & ~trunc
    out = {}
    for tag, tf in (("a", None), ("b", terminated_fn)):
        buf = HindsightExperienceReplayBuffer(cap, goal_dim, reward_fn, tf)
        for i in range(total):
            buf.push(state=torch.from_numpy(st[i].copy()), action=torch.tensor(int(ac[i])), reward=float(rw[i]), terminated=bool(term[i]),
                     truncated=bool(trunc[i]), curr_available_actions=space, next_state=torch.from_numpy(ns[i].copy()),
                     next_available_actions=space, max_number_actions=n_act)
        mem = list(buf.memory)
        out[f"state_{tag}"] = np.stack([t.state.reshape(-1).numpy() for t in mem])
        out[f"next_state_{tag}"] = np.stack([t.next_state.reshape(-1).numpy() for t in mem])
        out[f"reward_{tag}"] = np.asarray([float(t.reward) for t in mem], dtype=np.float32)
        out[f"action_{tag}"] = np.asarray([int(t.action.reshape(-1)[0]) for t in mem], dtype=np.int64)
        out[f"terminated_{tag}"] = np.asarray([bool(t.terminated) for t in mem])
        out[f"truncated_{tag}"] = np.asarray([bool(t.truncated) for t in mem])
    np.savez_compressed(os.path.join(GOLDEN, "her_small.npz"), goal_dim=goal_dim, n_act=n_act, capacity=cap, state=st, next_state=ns, reward=rw,
                        action=ac, terminated=term, truncated=trunc, **out)
    print("her_small.npz:", len(out["reward_a"]), "stored transitions")


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "her":
        gen_her_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "td3":
        gen_td3_golden("td3")
        gen_td3_golden("ddpg")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sac":
        gen_sac_golden()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ppo":
        gen_ppo_gae_golden()
        gen_ppo_learn_golden()
        sys.exit(0)
    gen_random_sample_vectors()
    gen_ppo_gae_golden()
    gen_ppo_learn_golden()
    gen_sac_golden()
    # small everything; ring wraps (n_push > capacity); pool branch (n=48 <= 85)
    run_dqn_case("dqn_tiny", obs=8, n_act=4, hidden=(16, 16), capacity=48, n_push=70, batch=16,
                 rounds=12, target_update_freq=5, tau=0.75, double=False, seed=11, data_seed=101,
                 snap_rounds=(1, 2, 4, 5, 10))
    run_dqn_case("ddqn_tiny", obs=8, n_act=4, hidden=(16, 16), capacity=48, n_push=70, batch=16,
                 rounds=12, target_update_freq=5, tau=0.75, double=True, seed=12, data_seed=102,
                 snap_rounds=(1, 2, 4, 5, 10))
    # dynamic action space: per-transition next_available_actions subsets + masks
    run_dqn_case("dqn_dynamic", obs=6, n_act=5, hidden=(16, 8), capacity=64, n_push=64, batch=32,
                 rounds=8, target_update_freq=3, tau=0.5, double=False, seed=13, data_seed=103,
                 dynamic=True, snap_rounds=(1, 3, 8))
    run_dqn_case("ddqn_dynamic", obs=6, n_act=5, hidden=(16, 8), capacity=64, n_push=64, batch=32,
                 rounds=8, target_update_freq=3, tau=0.5, double=True, seed=14, data_seed=104,
                 dynamic=True, snap_rounds=(1, 3, 8))
    # cfg1 shape (CartPole: obs=4, A=2, [64,64], B=32), two learn() calls
    run_dqn_case("dqn_cfg1", obs=4, n_act=2, hidden=(64, 64), capacity=512, n_push=400, batch=32,
                 rounds=10, learn_calls=2, target_update_freq=10, tau=0.75, double=False, seed=21,
                 data_seed=201, snap_rounds=(1, 10, 20))
    # cfg2 network shape (obs=128, A=16, [64,64], B=256) on a small buffer: pool branch (n<=1045)
    run_dqn_case("dqn_cfg2_pool", obs=128, n_act=16, hidden=(64, 64), capacity=600, n_push=600,
                 batch=256, rounds=20, target_update_freq=10, tau=0.75, double=False, seed=1234,
                 data_seed=4321, snap_rounds=(1, 9, 10, 20))
    # set branch (n = 1100 > 1045) with B=256; narrower obs to keep the fixture small
    run_dqn_case("ddqn_setbranch", obs=32, n_act=16, hidden=(64, 64), capacity=1100, n_push=1500,
                 batch=256, rounds=12, target_update_freq=10, tau=0.75, double=True, seed=77,
                 data_seed=770, snap_rounds=(1, 10, 12))
    # batch larger than buffer -> learn() clamps batch to len(buffer) (policy_learner.py:176-179)
    run_dqn_case("dqn_short_buffer", obs=8, n_act=4, hidden=(16, 16), capacity=100, n_push=10,
                 batch=32, rounds=3, target_update_freq=2, tau=0.75, double=False, seed=5,
                 data_seed=55, snap_rounds=(1, 3))
