"""Prioritized replay.  The reference has none (parity unpinned by the reference): the CUDA
sum tree / sampler / weighted learner are checked against oracle/per_oracle.py (bit-exact trees and indices
on identical leaves) and oracle/pearl_oracle.py (weighted DoubleDQN / DQN step, 1e-4)."""
import numpy as np
import pytest
import torch

from oracle.per_oracle import PerOracle, philox4x32_10


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_oracle_tree_invariants():
    o = PerOracle(37, seed=5)
    o.push(np.arange(37))
    rng = np.random.default_rng(0)
    o.set_leaves(rng.choice(37, 20, replace=False), rng.random(20).astype(np.float32) + 0.01)
    leaves = o.sum[o.C2:o.C2 + 37]
    assert abs(float(o.sum[1]) - float(leaves.astype(np.float64).sum())) < 1e-4
    assert o.min[1] == leaves.min()
    slots, w = o.sample(16, 3)
    assert slots.min() >= 0 and slots.max() < 37 and np.all(np.diff(slots) >= 0) and w.max() <= 1.0 + 1e-6


class _Space:
    def __init__(self, n):
        self.n = n
        self.actions = [torch.tensor([i]) for i in range(n)]

    @property
    def actions_batch(self):
        return torch.stack(self.actions)


@pytest.mark.gpu
def test_tree_and_sampler_match_oracle_bit_exactly():
    import pearl_b200
    from oracle.synth import make_transitions
    cap, n, obs, A, B = 3000, 4100, 8, 4, 64            # ring wraps: 1100 slots overwritten
    d = make_transitions(n, obs, A, seed=1)
    buf = pearl_b200.B200PrioritizedReplayBuffer(cap, alpha=0.6, beta=0.4, eps=1e-6, seed=0x1234567890ABCDEF & 0xFFFFFFFFFFFF)
    orc = PerOracle(cap, 0.6, 0.4, 1e-6, seed=0x1234567890ABCDEF & 0xFFFFFFFFFFFF)
    t = lambda k, a, b: torch.from_numpy(d[k][a:b])
    pos = 0
    for a, b in ((0, 1000), (1000, 1001), (1001, 4100)):
        buf.push_batch(t("state", a, b), t("action", a, b), t("reward", a, b), t("next_state", a, b), t("terminated", a, b),
                       t("truncated", a, b), max_number_actions=A)
        orc.push((np.arange(a, b) % cap))
    g = torch.Generator().manual_seed(3)
    for step in range(6):
        slots, w = buf.sample_prioritized(B)
        so, wo = orc.sample(B, step)
        assert slots.cpu().tolist() == so.tolist(), f"draw {step} differs"
        np.testing.assert_allclose(w.cpu().numpy(), wo, rtol=2e-6)
        td = torch.randn(B, generator=g) * (step + 1)
        pr = buf.update_priorities(slots, td).cpu().numpy()
        np.testing.assert_allclose(pr, orc.priority_of(td.numpy()), rtol=2e-6)
        orc.set_leaves(so, pr)                          # identical leaves on both sides from here on
        assert np.array_equal(buf.sum_tree.cpu().numpy()[1:], orc.sum[1:]), "sum tree differs"
        assert np.array_equal(buf.min_tree.cpu().numpy()[1:], orc.min[1:]), "min tree differs"
    assert float(buf._max_priority.cpu()) == float(orc.max_priority)
    b = buf.sample(B)
    assert b.weight is not None and b.weight.shape == (B,) and b.state.shape == (B, obs)


@pytest.mark.gpu
@pytest.mark.parametrize("double", [False, True])
def test_prioritized_learn_matches_weighted_oracle(double):
    import pearl_b200
    from oracle.pearl_oracle import OracleDQN, flat
    from oracle.synth import make_transitions
    cap, obs, A, B, hidden = 2000, 16, 4, 64, (32, 32)
    d = make_transitions(cap, obs, A, seed=8)
    buf = pearl_b200.B200PrioritizedReplayBuffer(cap, seed=77)
    buf.push_batch(*(torch.from_numpy(d[k]) for k in ("state", "action", "reward", "next_state", "terminated", "truncated")),
                   max_number_actions=A)
    orc_tree = PerOracle(cap, seed=77)
    orc_tree.push(np.arange(cap))
    torch.manual_seed(2)
    cls = pearl_b200.B200DoubleDQN if double else pearl_b200.B200DeepQLearning
    learner = cls(state_dim=obs, action_space=_Space(A), hidden_dims=list(hidden), training_rounds=1, batch_size=B,
                  target_update_freq=3, soft_update_tau=0.5,
                  action_representation_module=pearl_b200.OneHotActionTensorRepresentationModule(A)).to("cuda")
    orc = OracleDQN(obs, A, hidden, batch_size=B, target_update_freq=3, tau=0.5, double=double,
                    init_q=flat(learner._Q).cpu(), init_q_target=flat(learner._Q_target).cpu())
    eye = torch.eye(A).unsqueeze(0).expand(B, A, A)
    for r in range(8):
        rep = learner.learn(buf, trace=True)
        slots = rep["slots"][0].cpu().numpy()
        so, wo = orc_tree.sample(B, r)
        assert slots.tolist() == so.tolist(), f"round {r}: prioritized draw differs"
        np.testing.assert_allclose(rep["weight"][0].cpu().numpy(), wo, rtol=2e-6)
        orc.training_steps += 1
        b = dict(state=torch.from_numpy(d["state"][so]), action=orc._one_hot(torch.from_numpy(d["action"][so])),
                 reward=torch.from_numpy(d["reward"][so]), terminated=torch.from_numpy(d["terminated"][so]),
                 next_state=torch.from_numpy(d["next_state"][so]), next_available_actions=eye,
                 next_unavailable_actions_mask=torch.zeros((B, A), dtype=torch.bool), weight=torch.from_numpy(wo))
        orc.trace = {"q": [], "y": []}
        loss = orc.learn_batch(b)
        np.testing.assert_allclose(rep["q"][0].cpu().numpy(), orc.trace["q"][0].numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(rep["y"][0].cpu().numpy(), orc.trace["y"][0].numpy(), rtol=1e-4, atol=1e-6)
        assert abs(rep["loss"][0] - loss) <= 1e-4 * abs(loss) + 1e-6
        # priorities the kernel wrote (= leaves at the sampled slots) vs (|q-y| + eps)^alpha; then keep trees identical
        leaves = buf.sum_tree.cpu().numpy()[orc_tree.C2 + so]
        td = np.abs(rep["q"][0].cpu().numpy() - rep["y"][0].cpu().numpy())
        np.testing.assert_allclose(leaves, orc_tree.priority_of(td), rtol=2e-6)
        orc_tree.set_leaves(so, leaves)
        assert np.array_equal(buf.sum_tree.cpu().numpy()[1:], orc_tree.sum[1:])
    np.testing.assert_allclose(learner.flat_parameters.cpu().numpy(), flat(orc.Q).numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(learner.flat_target_parameters.cpu().numpy(), flat(orc.Qt).numpy(), rtol=1e-4, atol=1e-6)
