"""PPO preprocessing on the GPU: generalized advantage estimation and truncated lambda returns
(`ProximalPolicyOptimization.preprocess_replay_buffer`, pearl/policy_learners/
sequential_decision_making/ppo.py:201-293) for a whole rollout in one kernel launch."""
from __future__ import annotations

import torch

from . import _lib
from .replay_buffer import _stream_ptr


def gae_and_lambda_returns(values: torch.Tensor, last_next_value: float, reward: torch.Tensor,
                           terminated: torch.Tensor, truncated: torch.Tensor, discount_factor: float,
                           trace_decay_param: float):
    """values[i] = critic(state_i) in TIME order (0 = oldest), CUDA tensors.
    Returns (gae, lam_return), each [n] fp32, bit-identical to the reference's newest->oldest loop."""
    if not values.is_cuda:
        raise RuntimeError("pearl_b200 has no CPU path: the rollout must live on a CUDA device")
    dev = values.device
    lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
    n = values.numel()
    v = values.reshape(n).to(torch.float32).contiguous()
    r = reward.reshape(n).to(device=dev, dtype=torch.float32).contiguous()
    te = terminated.reshape(n).to(device=dev, dtype=torch.uint8).contiguous()
    tr = truncated.reshape(n).to(device=dev, dtype=torch.uint8).contiguous()
    gae = torch.empty(n, dtype=torch.float32, device=dev)
    lam = torch.empty(n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.prl_ppo_gae(n, _lib.ptr(v), float(last_next_value), _lib.ptr(r), _lib.ptr(te), _lib.ptr(tr),
                                   float(discount_factor), float(trace_decay_param), _lib.ptr(gae), _lib.ptr(lam),
                                   _stream_ptr(dev)))
    return gae, lam
