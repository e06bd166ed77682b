"""Developer tool: time the SAC learner (cfg3 shape) on one GPU."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import sys
import time

import torch

import pearl_b200

import os
obs, act, n, B = [int(x) for x in os.environ.get("SAC_SHAPE", "17,6,100000,256").split(",")]
R = int(sys.argv[1]) if len(sys.argv) > 1 else 200
buf = pearl_b200.B200ReplayBuffer(n, rng="device")
buf.is_action_continuous = True
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
buf.push_batch(rn(n, obs), rn(n, act).clamp(-1, 1), rn(n), rn(n, obs), torch.zeros(n, dtype=torch.bool, device="cuda"),
               torch.zeros(n, dtype=torch.bool, device="cuda"))
buf.seed(1)
pl = pearl_b200.B200ContinuousSoftActorCritic(state_dim=obs, low=[-1.0] * act, high=[1.0] * act, actor_hidden_dims=[256, 256],
                                              critic_hidden_dims=[256, 256], training_rounds=R, batch_size=B, seed=3)
pl.use_cuda_graph = os.environ.get('SAC_GRAPH', '1') == '1'
pl.learn(buf)
torch.cuda.synchronize()
t0 = time.perf_counter()
rep = pl.learn(buf)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"SAC cfg3: {R} steps in {dt*1e3:.1f} ms -> {dt/R*1e6:.1f} us/step, {R/dt:.0f} steps/s; "
      f"actor_loss {rep['actor_loss'][-1]:.4f} critic_loss {rep['critic_loss'][-1]:.4f} alpha {pl.entropy_coef:.4f}")
