"""Base classes for the drop-in plugins.

When facebookresearch/Pearl is importable, `B200ReplayBuffer` subclasses
`pearl.replay_buffers.replay_buffer.ReplayBuffer` and the learners subclass
`pearl...DeepQLearning` / `DoubleDQN`, so `pearl.pearl_agent.PearlAgent` accepts
them unchanged.  When it is not installed (e.g. the GPU test box) equivalent
stand-alone bases with the same attribute names are used; the CUDA path is
identical in both cases.  Nothing here computes anything.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass, fields
from typing import Any, Optional

import torch
from torch import nn

try:  # pragma: no cover - depends on the environment
    from pearl.action_representation_modules.one_hot_action_representation_module import (
        OneHotActionTensorRepresentationModule,
    )
    from pearl.policy_learners.sequential_decision_making.deep_q_learning import (
        DeepQLearning as _RefDeepQLearning,
    )
    from pearl.policy_learners.sequential_decision_making.double_dqn import DoubleDQN as _RefDoubleDQN
    from pearl.replay_buffers.replay_buffer import ReplayBuffer
    from pearl.replay_buffers.transition import TransitionBatch

    HAVE_PEARL = True
except Exception:  # ModuleNotFoundError (pearl or gymnasium missing)
    HAVE_PEARL = False

    class ReplayBuffer(ABC):  # mirrors pearl/replay_buffers/replay_buffer.py:18-91
        def __init__(self) -> None:
            super().__init__()
            self._is_action_continuous: bool = False
            self._has_cost_available: bool = False

        @property
        @abstractmethod
        def device_for_batches(self) -> torch.device: ...

        @abstractmethod
        def push(self, state, action, reward, terminated, truncated, curr_available_actions=None,
                 next_state=None, next_available_actions=None, max_number_actions=None,
                 cost=None) -> None: ...

        @abstractmethod
        def sample(self, batch_size: int): ...

        @abstractmethod
        def clear(self) -> None: ...

        @abstractmethod
        def __len__(self) -> int: ...

        @property
        def is_action_continuous(self) -> bool:
            return self._is_action_continuous

        @is_action_continuous.setter
        def is_action_continuous(self, value: bool) -> None:
            self._is_action_continuous = value

    @dataclass
    class TransitionBatch:  # field names of pearl/replay_buffers/transition.py:89-130
        state: torch.Tensor
        action: torch.Tensor
        reward: torch.Tensor
        terminated: Optional[torch.Tensor] = None
        truncated: Optional[torch.Tensor] = None
        next_state: Optional[torch.Tensor] = None
        next_action: Optional[torch.Tensor] = None
        curr_available_actions: Optional[torch.Tensor] = None
        curr_unavailable_actions_mask: Optional[torch.Tensor] = None
        next_available_actions: Optional[torch.Tensor] = None
        next_unavailable_actions_mask: Optional[torch.Tensor] = None
        weight: Optional[torch.Tensor] = None
        time_diff: Optional[torch.Tensor] = None
        cost: Optional[torch.Tensor] = None

        def __post_init__(self) -> None:
            n = self.reward.shape[0]
            if self.terminated is None:
                self.terminated = torch.ones(n, dtype=torch.bool, device=self.reward.device)
            if self.truncated is None:
                self.truncated = torch.zeros(n, dtype=torch.bool, device=self.reward.device)

        def to(self, device):
            for f in fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    setattr(self, f.name, torch.as_tensor(v, device=device))
            return self

        @property
        def device(self) -> torch.device:
            return self.state.device

        def __len__(self) -> int:
            return self.reward.shape[0]

    class OneHotActionTensorRepresentationModule(nn.Module):
        def __init__(self, max_number_actions: int) -> None:
            super().__init__()
            self._max_number_actions = max_number_actions

        def forward(self, x: torch.Tensor) -> torch.Tensor:
            if x.dim() == 1:
                x = x.unsqueeze(-1)
            return torch.nn.functional.one_hot(x.long(), self._max_number_actions).squeeze(-2).float()

        @property
        def max_number_actions(self) -> int:
            return self._max_number_actions

        @property
        def representation_dim(self) -> int:
            return self._max_number_actions

    class _QNet(nn.Module):
        """Same module tree (hence state_dict keys) as VanillaQValueNetwork built by
        mlp_block: `_model.{i}.0` = Linear (pearl/neural_networks/common/utils.py:75-152)."""

        def __init__(self, state_dim: int, action_dim: int, hidden_dims) -> None:
            super().__init__()
            dims = [state_dim + action_dim] + list(hidden_dims) + [1]
            layers = [nn.Sequential(nn.Linear(dims[i], dims[i + 1]), nn.ReLU()) for i in range(len(dims) - 2)]
            layers.append(nn.Sequential(nn.Linear(dims[-2], dims[-1])))
            self._model = nn.Sequential(*layers)
            self._state_dim, self._action_dim = state_dim, action_dim

    class _RefDeepQLearning(nn.Module):
        """Attribute-compatible stand-in for DeepQLearning's constructor
        (deep_q_learning.py:40-59, deep_td_learning.py:60-185)."""

        def __init__(self, action_space=None, hidden_dims=None, exploration_module=None,
                     learning_rate: float = 0.001, discount_factor: float = 0.99,
                     training_rounds: int = 10, batch_size: int = 128, target_update_freq: int = 10,
                     soft_update_tau: float = 0.75, is_conservative: bool = False,
                     conservative_alpha: Optional[float] = 2.0, state_dim: Optional[int] = None,
                     network_type=None, action_representation_module=None, network_instance=None,
                     optimizer=None, **kwargs: Any) -> None:
            super().__init__()
            import copy
            assert state_dim is not None and hidden_dims is not None
            assert action_representation_module is not None
            self._action_space = action_space
            self._training_rounds, self._batch_size = training_rounds, batch_size
            self._training_steps = 0
            self._learning_rate, self._discount_factor = learning_rate, discount_factor
            self._target_update_freq, self._soft_update_tau = target_update_freq, soft_update_tau
            self._is_conservative = is_conservative
            self._is_action_continuous = False
            self.on_policy = False
            self.exploration_module = exploration_module
            self._action_representation_module = action_representation_module
            self._Q = network_instance if network_instance is not None else _QNet(
                state_dim, action_representation_module.representation_dim, hidden_dims)
            self._Q_target = copy.deepcopy(self._Q)
            self._optimizer = optimizer if optimizer is not None else torch.optim.AdamW(
                self._Q.parameters(), lr=learning_rate, amsgrad=True)

        @property
        def action_representation_module(self):
            return self._action_representation_module

        @property
        def requires_tensors(self) -> bool:
            return True

        @property
        def batch_size(self) -> int:
            return self._batch_size

        @property
        def optimizer(self):
            return self._optimizer

        def set_history_summarization_module(self, value) -> None:
            self._history_summarization_module = value

        def reset(self, action_space) -> None:
            self._action_space = action_space

    class _RefDoubleDQN(_RefDeepQLearning):
        pass
