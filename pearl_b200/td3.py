"""B200TD3 / B200DeepDeterministicPolicyGradient — the learner side of Pearl's TD3
(pearl/policy_learners/sequential_decision_making/td3.py:43-202) and DeepDeterministicPolicyGradient (ddpg.py:41-157, on
actor_critic_base.py:309-366) on a B200: `learn(replay_buffer)` runs `training_rounds` x (sample -> [delayed] actor step ->
twin-critic step -> [delayed] soft target updates) on the GPU through `prl_td3_learn` (include/pearl_b200.h).  Same constructor
argument names and the same report keys (`actor_loss`, `critic_loss`) as the reference.  PyTorch holds the flat parameter
vectors and draws the target-policy noise (the reference's `torch.normal`); no math happens in Python.  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Any, Iterable, Optional

import torch

from . import _lib
from .replay_buffer import B200ReplayBuffer, _stream_ptr
from .sac import _bounds


class B200TD3:
    _default_freq, _default_noise, _default_clip = 2, 0.2, 0.5

    def __init__(self, state_dim: int, action_space: Any = None, actor_hidden_dims: Optional[Iterable[int]] = None,
                 critic_hidden_dims: Optional[Iterable[int]] = None, actor_learning_rate: float = 1e-3,
                 critic_learning_rate: float = 1e-3, actor_soft_update_tau: float = 0.005, critic_soft_update_tau: float = 0.005,
                 discount_factor: float = 0.99, training_rounds: int = 1, batch_size: int = 256,
                 actor_update_freq: Optional[int] = None, actor_update_noise: Optional[float] = None,
                 actor_update_noise_clip: Optional[float] = None, *, low=None, high=None,
                 device: Optional[torch.device | str | int] = None, max_rounds_per_call: int = 1024, seed: Optional[int] = None) -> None:
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self._device.index is None:
            self._device = torch.device("cuda", torch.cuda.current_device())
        self._lib = _lib.init(self._device.index)
        actor_hidden_dims, critic_hidden_dims = list(actor_hidden_dims or []), list(critic_hidden_dims or [])
        if len(actor_hidden_dims) != 2 or len(critic_hidden_dims) != 2:
            raise NotImplementedError("the CUDA TD3 / DDPG learner is built for two hidden layers in the actor and in each critic")
        self._state_dim = int(state_dim)
        self._low, self._high = _bounds(action_space, low, high, self._device)
        self._action_dim = int(self._low.numel())
        self._actor_hidden_dims, self._critic_hidden_dims = actor_hidden_dims, critic_hidden_dims
        self._actor_learning_rate, self._critic_learning_rate = float(actor_learning_rate), float(critic_learning_rate)
        self._actor_soft_update_tau, self._critic_soft_update_tau = float(actor_soft_update_tau), float(critic_soft_update_tau)
        self._discount_factor = float(discount_factor)
        self._training_rounds, self._batch_size = int(training_rounds), int(batch_size)
        self._actor_update_freq = int(self._default_freq if actor_update_freq is None else actor_update_freq)
        self._actor_update_noise = float(self._default_noise if actor_update_noise is None else actor_update_noise)
        self._actor_update_noise_clip = float(self._default_clip if actor_update_noise_clip is None else actor_update_noise_clip)
        self._max_rounds = max(int(max_rounds_per_call), 1)
        self._training_steps = 0
        self.use_cuda_graph = True
        self._handle = C.c_void_p(0)
        self._bound_batch = 0
        self._gen = torch.Generator(device=self._device)
        if seed is not None:
            self._gen.manual_seed(int(seed))
        cfg = self._cfg(1)
        pa, pc = int(self._lib.prl_td3_actor_param_count(C.byref(cfg))), int(self._lib.prl_td3_critic_param_count(C.byref(cfg)))
        dev, f32 = self._device, torch.float32
        self.actor_params = torch.empty(pa, dtype=f32, device=dev)
        self.critic_params = torch.empty(2 * pc, dtype=f32, device=dev)
        self._init_like_reference()
        self.actor_target_params = self.actor_params.clone()
        self.critic_target_params = self.critic_params.clone()
        self._actor_state = [torch.zeros(pa, dtype=f32, device=dev) for _ in range(3)]      # exp_avg, exp_avg_sq, max_exp_avg_sq
        self._critic_state = [torch.zeros(2 * pc, dtype=f32, device=dev) for _ in range(3)]
        self._adam_steps = (0, 0)

    def _cfg(self, max_batch: int) -> _lib.Td3Cfg:
        return _lib.Td3Cfg(self._state_dim, self._action_dim, self._actor_hidden_dims[0], self._actor_hidden_dims[1],
                           self._critic_hidden_dims[0], self._critic_hidden_dims[1], self._actor_update_freq, max_batch, self._max_rounds,
                           self._actor_learning_rate, self._critic_learning_rate, 0.9, 0.999, 1e-8, 0.01, self._discount_factor,
                           self._actor_soft_update_tau, self._critic_soft_update_tau, self._actor_update_noise_clip)

    def _shapes(self):
        O, A, (h1, h2), (c1, c2) = self._state_dim, self._action_dim, self._actor_hidden_dims, self._critic_hidden_dims
        return [(h1, O), (h1,), (h2, h1), (h2,), (A, h2), (A,)], [(c1, O + A), (c1,), (c2, c1), (c2,), (1, c2), (1,)]

    def _init_like_reference(self) -> None:
        """Xavier-uniform weights, biases 0.01 (neural_networks/common/utils.py:201-205, actor_critic_base.py:154, twin_critic.py:36-60)."""
        def fill(vec, shapes):
            off = 0
            for shp in shapes:
                n = shp[0] * (shp[1] if len(shp) == 2 else 1)
                if len(shp) == 2:
                    bound = (6.0 / (shp[0] + shp[1])) ** 0.5
                    vec[off:off + n].uniform_(-bound, bound, generator=self._gen)
                else:
                    vec[off:off + n].fill_(0.01)
                off += n
            assert off == vec.numel()
        sa, sc = self._shapes()
        fill(self.actor_params, sa)
        pc = self.critic_params.numel() // 2
        fill(self.critic_params[:pc], sc)
        fill(self.critic_params[pc:], sc)

    def load_parameters(self, actor, q1, q2, actor_target=None, q1_target=None, q2_target=None) -> None:
        """Flat fp32 vectors in `torch.nn.Module.parameters()` order of the reference networks."""
        t = lambda x: torch.as_tensor(x, dtype=torch.float32).reshape(-1).to(self._device)  # noqa: E731
        pc = self.critic_params.numel() // 2
        self.actor_params.copy_(t(actor))
        self.actor_target_params.copy_(t(actor if actor_target is None else actor_target))
        self.critic_params[:pc].copy_(t(q1)); self.critic_params[pc:].copy_(t(q2))
        self.critic_target_params[:pc].copy_(t(q1 if q1_target is None else q1_target))
        self.critic_target_params[pc:].copy_(t(q2 if q2_target is None else q2_target))

    @property
    def batch_size(self) -> int:
        return self._batch_size

    @property
    def training_rounds(self) -> int:
        return self._training_rounds

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                self._lib.prl_td3_destroy(self._handle)
                self._handle = C.c_void_p(0)
        except Exception:
            pass

    def _bind(self, need_batch: int) -> None:
        if self._handle.value and need_batch <= self._bound_batch:
            return
        if self._handle.value:
            self._adam_steps = (int(self._lib.prl_td3_actor_adam_step(self._handle)), int(self._lib.prl_td3_critic_adam_step(self._handle)))
            self._lib.prl_td3_destroy(self._handle)
            self._handle = C.c_void_p(0)
        cfg = self._cfg(max(need_batch, self._batch_size if self._batch_size > 0 else need_batch))
        self._workspace = torch.empty(int(self._lib.prl_td3_workspace_bytes(C.byref(cfg))), dtype=torch.uint8, device=self._device)
        h = C.c_void_p(0)
        p = _lib.ptr
        with torch.cuda.device(self._device):
            _lib.check(self._lib.prl_td3_create(
                C.byref(h), C.byref(cfg), p(self.actor_params), p(self._actor_state[0]), p(self._actor_state[1]), p(self._actor_state[2]),
                p(self.actor_target_params), p(self.critic_params), p(self._critic_state[0]), p(self._critic_state[1]),
                p(self._critic_state[2]), p(self.critic_target_params), p(self._low), p(self._high), self._adam_steps[0],
                self._adam_steps[1], p(self._workspace)))
        self._handle, self._bound_batch = h, cfg.max_batch

    # ------------------------------------------------------------------ PolicyLearner.learn (policy_learner.py:162-204)
    def learn(self, replay_buffer: B200ReplayBuffer, noise: Optional[torch.Tensor] = None, trace: Optional[dict] = None) -> dict:
        if not isinstance(replay_buffer, B200ReplayBuffer):
            raise TypeError(f"{type(self).__name__} learns from a B200ReplayBuffer (GPU-resident ring)")
        if len(replay_buffer) == 0:
            return {}
        if not replay_buffer.is_action_continuous:
            raise ValueError("TD3 / DDPG need a replay buffer with is_action_continuous=True")
        B = len(replay_buffer) if (self._batch_size == -1 or len(replay_buffer) < self._batch_size) else self._batch_size
        self._bind(B)
        R, A, dev = self._training_rounds, self._action_dim, self._device
        report = {"actor_loss": [], "critic_loss": []}
        idx_all = []
        done = 0
        while done < R:
            r = min(self._max_rounds, R - done)
            nz = None
            if self._actor_update_noise > 0.0:
                if noise is not None:
                    nz = noise[done:done + r].to(device=dev, dtype=torch.float32).contiguous()
                    if tuple(nz.shape) != (r, B, A):
                        raise ValueError(f"noise must be [rounds, {B}, {A}]")
                else:       # torch.normal(mean=0, std=actor_update_noise, size=next_action.size())   (td3.py:155-160)
                    nz = torch.randn((r, B, A), dtype=torch.float32, device=dev, generator=self._gen) * self._actor_update_noise
            out = torch.empty((2, r), dtype=torch.float32, device=dev)
            idx = torch.empty((r, B), dtype=torch.int32, device=dev) if trace is not None else None
            replay_buffer._rng_push()
            with torch.cuda.device(dev):
                _lib.check(self._lib.prl_td3_set_graph(self._handle, int(self.use_cuda_graph)))
                _lib.check(self._lib.prl_td3_learn(self._handle, replay_buffer.handle, r, B, int(self._training_steps), _lib.ptr(nz),
                                                   _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(idx) if idx is not None else None,
                                                   _stream_ptr(dev)))
            replay_buffer._rng_pull()
            host = out.cpu()
            report["actor_loss"] += host[0].tolist()
            report["critic_loss"] += host[1].tolist()
            if idx is not None:
                idx_all.append(idx.cpu())
            self._training_steps += r
            done += r
        if trace is not None:
            trace["idx"] = torch.cat(idx_all)
        return report


class B200DeepDeterministicPolicyGradient(B200TD3):
    """DDPG = the same step with an actor update every round and no target-policy noise (ddpg.py:41-157)."""
    _default_freq, _default_noise, _default_clip = 1, 0.0, 0.0

    def __init__(self, *args, **kwargs) -> None:
        for k in ("actor_update_freq", "actor_update_noise", "actor_update_noise_clip"):
            if kwargs.get(k) is not None:
                raise TypeError(f"DeepDeterministicPolicyGradient has no `{k}` (use B200TD3)")
        super().__init__(*args, **kwargs)
