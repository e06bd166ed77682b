"""PPO preprocessing (GAE + truncated lambda returns): oracle vs the reference recording (CPU) and the
CUDA kernel vs both (GPU).  The recording comes from the reference's own preprocess_replay_buffer
(ppo.py:201-293) on a 700-step rollout with terminated / truncated / adjacent episode ends; `kat_*` is the
closed form of the reference's unit test (test_ppo.py:48-115: gamma 0.6, lambda 0.5, rewards 4, 6, 5)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.ppo_oracle import gae_reference_loop

FX = np.load(os.path.join(GOLDEN, "ppo_gae.npz"))


def test_oracle_reproduces_reference_recording_bit_exactly():
    g, l = gae_reference_loop(torch.from_numpy(FX["values"]), float(FX["last_next_value"]), torch.from_numpy(FX["reward"]),
                              torch.from_numpy(FX["terminated"]), torch.from_numpy(FX["truncated"]), float(FX["gamma"]),
                              float(FX["lam"]))
    assert np.array_equal(g.numpy(), FX["gae"])
    assert np.array_equal(l.numpy(), FX["lam_return"])


def test_oracle_closed_form_of_the_reference_unit_test():
    v = torch.from_numpy(FX["kat_v"])
    g, l = gae_reference_loop(v[:3], float(v[3]), torch.from_numpy(FX["kat_reward"]), torch.zeros(3, dtype=torch.bool),
                              torch.zeros(3, dtype=torch.bool), 0.6, 0.5)
    np.testing.assert_allclose(g.numpy(), FX["kat_gae"], rtol=1e-6)
    np.testing.assert_allclose(l.numpy(), FX["kat_gae"] + FX["kat_v"][:3], rtol=1e-6)


@pytest.mark.gpu
def test_kernel_matches_reference_recording_bit_exactly():
    from pearl_b200.ppo import gae_and_lambda_returns
    c = lambda k: torch.from_numpy(FX[k]).cuda()
    g, l = gae_and_lambda_returns(c("values"), float(FX["last_next_value"]), c("reward"), c("terminated"), c("truncated"),
                                  float(FX["gamma"]), float(FX["lam"]))
    assert np.array_equal(g.cpu().numpy(), FX["gae"])
    assert np.array_equal(l.cpu().numpy(), FX["lam_return"])
    v = torch.from_numpy(FX["kat_v"]).cuda()
    g, l = gae_and_lambda_returns(v[:3], float(v[3]), torch.from_numpy(FX["kat_reward"]).cuda(), torch.zeros(3, dtype=torch.bool).cuda(),
                                  torch.zeros(3, dtype=torch.bool).cuda(), 0.6, 0.5)
    np.testing.assert_allclose(g.cpu().numpy(), FX["kat_gae"], rtol=1e-6)


@pytest.mark.gpu
def test_kernel_full_size_rollout_matches_oracle():
    """BASELINE cfg4 shape: 65 536 consecutive steps, episodes end every 500 steps (+ random truncations)."""
    from pearl_b200.ppo import gae_and_lambda_returns
    n = 65_536
    g_ = torch.Generator().manual_seed(4)
    values = torch.randn(n, generator=g_)
    reward = torch.randn(n, generator=g_)
    terminated = (torch.arange(n) % 500) == 499
    truncated = torch.rand(n, generator=g_) < 0.001
    want_g, want_l = gae_reference_loop(values, 0.25, reward, terminated, truncated, 0.99, 0.95)
    got_g, got_l = gae_and_lambda_returns(values.cuda(), 0.25, reward.cuda(), terminated.cuda(), truncated.cuda(), 0.99, 0.95)
    assert torch.equal(got_g.cpu(), want_g) and torch.equal(got_l.cpu(), want_l)
    # no episode boundary at all: one 65 536-long chain (worst case for the per-episode walk), still exact
    z = torch.zeros(n, dtype=torch.bool)
    want_g, _ = gae_reference_loop(values[:5000], 0.25, reward[:5000], z[:5000], z[:5000], 0.99, 0.95)
    got_g, _ = gae_and_lambda_returns(values[:5000].cuda(), 0.25, reward[:5000].cuda(), z[:5000].cuda(), z[:5000].cuda(), 0.99, 0.95)
    assert torch.equal(got_g.cpu(), want_g)
    # timing (informational): the reference's Python loop manages ~6.8 k transitions/s
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    args = (values.cuda(), 0.25, reward.cuda(), terminated.cuda(), truncated.cuda(), 0.99, 0.95)
    gae_and_lambda_returns(*args)
    e0.record()
    for _ in range(10):
        gae_and_lambda_returns(*args)
    e1.record()
    torch.cuda.synchronize()
    print(f"    GAE over {n} transitions: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per rollout")
