"""Host-contract tests on the GPU box: RNG stream hygiene, learner-group preconditions, optimizer / checkpoint
hand-over.  These pin the behaviours the round-1 code review flagged."""
import copy
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Space:
    def __init__(self, n):
        self.n = n
        self.actions = [torch.tensor([i]) for i in range(n)]

    @property
    def actions_batch(self):
        return torch.stack(self.actions)


def _buffer(n, obs, A, rng="device", seed=None):
    import pearl_b200
    from oracle.synth import make_transitions
    d = make_transitions(n, obs, A, seed=7)
    buf = pearl_b200.B200ReplayBuffer(n, rng=rng)
    buf.push_batch(*(torch.from_numpy(d[k]) for k in ("state", "action", "reward", "next_state", "terminated", "truncated")),
                   max_number_actions=A)
    if seed is not None:
        buf.seed(seed)
    return buf


def _learner(obs, A, rounds=6, B=32, **kw):
    import pearl_b200
    return pearl_b200.B200DeepQLearning(
        state_dim=obs, action_space=_Space(A), hidden_dims=[64, 64], training_rounds=rounds, batch_size=B,
        target_update_freq=4, soft_update_tau=0.5,
        action_representation_module=pearl_b200.OneHotActionTensorRepresentationModule(A), **kw).to("cuda")


def test_unseeded_device_rng_buffer_samples_and_streams_differ():
    """rng="device" without seed(): the private MT19937 stream is initialised (never all-zero, which would hang the
    set-branch sampler), distinct per buffer, and sample() returns distinct in-range indices."""
    n, k = 5000, 64                        # n > setsize(64) = 277: set branch
    a, b = _buffer(n, 4, 2), _buffer(n, 4, 2)
    sa, sb = a.get_rng_state(), b.get_rng_state()
    assert sa[:624].any() and sb[:624].any() and not np.array_equal(sa, sb)
    la, _ = a.sample_indices(k, rounds=3)
    la = la.cpu().numpy()
    assert la.min() >= 0 and la.max() < n
    for row in la:
        assert len(set(row.tolist())) == k
    assert len(a.sample(k)) == k


def test_all_zero_mt_state_is_rejected():
    buf = _buffer(100, 4, 2)
    with pytest.raises(ValueError):
        buf.set_rng_state(np.zeros(625, dtype=np.uint32))


def test_rng_pull_keeps_gauss_next():
    """The reference's random.sample never touches the cached second Gaussian of random.gauss()."""
    buf = _buffer(2000, 4, 2, rng="python")
    random.seed(5)
    random.gauss(0.0, 1.0)                 # leaves gauss_next cached
    g = random.getstate()[2]
    assert g is not None
    buf.sample(16)
    assert random.getstate()[2] == g


def test_group_rejects_python_rng_buffers():
    import pearl_b200
    bufs = [_buffer(2000, 8, 4, rng="python") for _ in range(2)]
    ls = [_learner(8, 4, B=128, engine="tc") for _ in range(2)]
    with pytest.raises(ValueError):
        pearl_b200.B200LearnerGroup(ls, bufs).learn()


def test_optimizer_load_state_dict_and_lr_change_are_picked_up():
    """Checkpoint resume: learner A trains, its optimizer state_dict is loaded into learner B (same weights); both
    then take identical further steps.  A later change of param_groups[0]['lr'] takes effect."""
    obs, A = 8, 4
    torch.manual_seed(3)
    la = _learner(obs, A)
    lb = _learner(obs, A)
    buf_a, buf_b = _buffer(3000, obs, A, seed=11), _buffer(3000, obs, A, seed=11)
    la.learn(buf_a)
    lb.learn(buf_b)                        # binds B (its own moments), then diverge B's state on purpose
    lb.learn(buf_b)
    # resume B from A's checkpoint: parameters, target, optimizer
    sd = copy.deepcopy(la.state_dict())
    osd = copy.deepcopy(la.optimizer.state_dict())
    lb.load_state_dict(sd)
    lb.optimizer.load_state_dict(osd)
    lb._training_steps = la._training_steps
    buf_b.set_rng_state(buf_a.get_rng_state())
    ra, rb = la.learn(buf_a), lb.learn(buf_b)
    np.testing.assert_allclose(rb["loss"], ra["loss"], rtol=1e-6)
    assert torch.equal(lb.flat_parameters, la.flat_parameters)
    assert lb.adam_state()["step"] == la.adam_state()["step"]
    assert torch.equal(lb.adam_state()["exp_avg_sq"], la.adam_state()["exp_avg_sq"])
    # the optimizer state exposed through torch is again a view of the flat vectors
    p0 = next(iter(lb._Q.parameters()))
    assert lb.optimizer.state[p0]["exp_avg"].data_ptr() == lb.adam_state()["exp_avg"].data_ptr()
    # lr = 0: AdamW only applies the (now also zero) decoupled decay -> parameters frozen
    before = lb.flat_parameters.clone()
    lb.optimizer.param_groups[0]["lr"] = 0.0
    lb.learn(buf_b)
    assert torch.equal(lb.flat_parameters, before)


def test_prioritized_push_after_dynamic_upgrade_gets_priorities():
    """A push that upgrades the storage to dynamic action sets on a wrapped ring: the new rows' leaves carry the
    max priority (they can be sampled), at the slots the rows were written to."""
    import pearl_b200
    from oracle.synth import make_transitions
    cap, obs, A = 64, 4, 4
    buf = pearl_b200.B200PrioritizedReplayBuffer(cap, seed=1)
    t = torch.from_numpy
    d = make_transitions(100, obs, A, seed=1)     # wraps: 100 > 64
    buf.push_batch(*(t(d[k]) for k in ("state", "action", "reward", "next_state", "terminated", "truncated")), max_number_actions=A)
    d2 = make_transitions(10, obs, A, seed=2, dynamic=True)
    buf.push_batch(*(t(d2[k]) for k in ("state", "action", "reward", "next_state", "terminated", "truncated")), max_number_actions=A,
                   next_available_ids=t(d2["next_avail_ids"].astype(np.uint8)), next_available_count=t(d2["next_avail_n"].astype(np.int32)))
    assert len(buf) == cap
    C2 = buf.sum_tree.numel() // 2
    leaves = buf.sum_tree[C2:C2 + cap].cpu().numpy()
    assert (leaves > 0).all()                     # every stored row, old and new, can be drawn
    np.testing.assert_allclose(float(buf.sum_tree[1]), leaves.astype(np.float64).sum(), rtol=1e-5)


def test_buffer_snapshot_round_trip_and_offline_loader():
    """f4: `state_dict()` / `load_state_dict()` of the device ring (contents in FIFO order + sampler stream), and the
    reference's offline-data format (list of transition dicts, offline_learning_and_evaluation.py:39-137) loaded in
    chunks — same ring as one `push()` per transition."""
    import io
    import pearl_b200
    from oracle.synth import make_transitions
    cap, obs, A = 300, 6, 4
    d = make_transitions(450, obs, A, seed=9)                 # wraps
    t = torch.from_numpy
    keys = ("state", "action", "reward", "next_state", "terminated", "truncated")
    a = pearl_b200.B200ReplayBuffer(cap, rng="device")
    a.push_batch(*(t(d[k]) for k in keys), max_number_actions=A)
    a.seed(21)
    a.sample(16)                                              # advance the stream before the snapshot
    blob = io.BytesIO()
    torch.save(a.state_dict(), blob)
    blob.seek(0)
    b = pearl_b200.B200ReplayBuffer(cap, rng="device")
    b.load_state_dict(torch.load(blob, weights_only=False))
    assert len(b) == len(a) == cap
    la, _ = a.sample_indices(32, rounds=3)
    lb, _ = b.sample_indices(32, rounds=3)
    assert torch.equal(la, lb)
    ba, bb = a._gather_logical(la[0]), b._gather_logical(lb[0])
    for k in ("state", "next_state", "reward", "action", "terminated"):
        assert torch.equal(ba[k], bb[k]), k

    class Discrete:                                           # what a gym offline data set stores
        def __init__(self, n):
            self.n = n
    rows = [dict(observation=t(d["state"][i]), action=int(d["action"][i]), reward=float(d["reward"][i]),
                 next_observation=t(d["next_state"][i]), curr_available_actions=Discrete(A), next_available_actions=Discrete(A),
                 done=bool(d["terminated"][i])) for i in range(450)]
    c = pearl_b200.B200ReplayBuffer(cap, rng="device")
    assert c.load_offline_data(rows, max_number_actions_if_discrete=A, chunk=128) == 450
    e = pearl_b200.B200ReplayBuffer(cap, rng="device")
    for r in rows[:40]:                                       # the per-transition path of the reference loader
        e.push(state=r["observation"], action=r["action"], reward=r["reward"], next_state=r["next_observation"],
               curr_available_actions=_Space(A), next_available_actions=_Space(A), terminated=r["done"], truncated=False,
               max_number_actions=A)
    assert len(c) == cap
    gc = c._gather_logical(torch.arange(cap, dtype=torch.int32, device=c.device))
    want = {k: t(d[k][150:]) for k in keys}
    assert torch.equal(gc["state"].cpu(), want["state"]) and torch.equal(gc["reward"].cpu(), want["reward"])
    assert torch.equal(gc["action"].cpu(), want["action"].long()) and torch.equal(gc["terminated"].cpu(), want["terminated"])
    ge = e._gather_logical(torch.arange(40, dtype=torch.int32, device=e.device))
    assert torch.equal(ge["state"].cpu(), t(d["state"][:40])) and torch.equal(ge["mask"].cpu(), torch.zeros((40, A), dtype=torch.bool))


def test_act_is_greedy_over_available_actions_and_explores_like_the_reference():
    """f2: `act()` (deep_td_learning.py:200-254): exploit = argmax over the AVAILABLE actions of the fused Q forward;
    with an epsilon-greedy module the draw consumes the global `random` stream exactly like the reference's module."""
    import pearl_b200
    obs, A = 10, 6
    torch.manual_seed(0)
    learner = _learner(obs, A)
    s = torch.randn(obs)
    q = learner.q_values(s.reshape(1, -1))[0].cpu()
    full = _Space(A)
    assert int(learner.act(s, full, exploit=True)) == int(q.argmax())
    sub = _Space(A)
    sub.actions = [torch.tensor([i]) for i in (1, 3, 4)]
    sub.n = 3
    assert int(learner.act(s, sub, exploit=True)) == (1, 3, 4)[int(q[[1, 3, 4]].argmax())]

    class EGreedy:                                            # epsilon_greedy_exploration.py:52-83
        def act(self, subjective_state, action_space, exploit_action, values=None, **kw):
            if random.random() < 0.5:
                return action_space.actions[random.randrange(action_space.n)]
            return exploit_action
    learner.exploration_module = EGreedy()
    random.seed(3)
    got = [int(learner.act(s, full)) for _ in range(20)]
    random.seed(3)
    want = []
    for _ in range(20):
        want.append(random.randrange(A) if random.random() < 0.5 else int(q.argmax()))
    assert got == want


@pytest.mark.parametrize("tag", ["a", "b"])
def test_hindsight_buffer_matches_the_reference_recording(tag):
    """f1: B200HindsightExperienceReplayBuffer against the contents of the reference's HindsightExperienceReplayBuffer
    (tests/golden/her_small.npz: three episodes, FIFO eviction at capacity 40; "b" = with a terminated_fn)."""
    import os
    import pearl_b200
    from conftest import GOLDEN
    fx = np.load(os.path.join(GOLDEN, "her_small.npz"))
    gd, A, cap = int(fx["goal_dim"]), int(fx["n_act"]), int(fx["capacity"])
    reward_fn = lambda s, a: float((s[:gd] - s[-gd:]).abs().sum() < 0.75) - 0.5          # noqa: E731 (as in oracle/gen_golden.py)
    terminated_fn = (lambda s, a: bool((s[:gd] - s[-gd:]).abs().sum() < 0.25)) if tag == "b" else None
    buf = pearl_b200.B200HindsightExperienceReplayBuffer(cap, gd, reward_fn, terminated_fn)
    t = torch.from_numpy
    for i in range(fx["state"].shape[0]):
        buf.push(state=t(fx["state"][i].copy()), action=torch.tensor(int(fx["action"][i])), reward=float(fx["reward"][i]),
                 terminated=bool(fx["terminated"][i]), truncated=bool(fx["truncated"][i]), curr_available_actions=_Space(A),
                 next_state=t(fx["next_state"][i].copy()), next_available_actions=_Space(A), max_number_actions=A)
    n = len(buf)
    assert n == fx[f"reward_{tag}"].shape[0]
    g = buf._gather_logical(torch.arange(n, dtype=torch.int32, device=buf.device))
    assert np.array_equal(g["state"].cpu().numpy(), fx[f"state_{tag}"]) and np.array_equal(g["next_state"].cpu().numpy(), fx[f"next_state_{tag}"])
    assert np.array_equal(g["reward"].cpu().numpy(), fx[f"reward_{tag}"]) and np.array_equal(g["action"].cpu().numpy(), fx[f"action_{tag}"])
    assert np.array_equal(g["terminated"].cpu().numpy(), fx[f"terminated_{tag}"]) and np.array_equal(g["truncated"].cpu().numpy(), fx[f"truncated_{tag}"])
