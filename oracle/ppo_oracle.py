"""oracle/ppo_oracle.py — CPU restatement of PPO's GAE preprocessing (TEST INFRASTRUCTURE ONLY).

Follows pearl/policy_learners/sequential_decision_making/ppo.py:271-293 literally (newest -> oldest,
same tensor expressions), on arrays in time order.  Pinned by tests/golden/ppo_gae.npz, recorded from
the real reference by oracle/gen_golden.py, and by the closed form of the reference's own unit test
(test/unit/with_pytorch/test_ppo.py:48-115).
"""
from __future__ import annotations

import torch


def gae_reference_loop(values, last_next_value, reward, terminated, truncated, discount_factor, trace_decay_param):
    n = values.shape[0]
    gae_out = torch.empty(n, dtype=torch.float32)
    lam_out = torch.empty(n, dtype=torch.float32)
    next_value = torch.as_tensor(last_next_value, dtype=torch.float32).reshape(1)
    gae = torch.tensor([0.0])
    for t in range(n - 1, -1, -1):
        term, trunc = terminated[t].reshape(1), truncated[t].reshape(1)
        td_error = reward[t].reshape(1) + discount_factor * next_value * (~term) - values[t].reshape(1)
        gae = td_error + discount_factor * trace_decay_param * (not (bool(term) or bool(trunc))) * gae
        gae_out[t] = gae[0]
        lam_out[t] = (gae + values[t].reshape(1))[0]
        next_value = values[t].reshape(1)
    return gae_out, lam_out
