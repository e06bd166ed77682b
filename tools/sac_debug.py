"""Developer tool: per-parameter-block error of the SAC learner vs the oracle after R rounds."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import random
import sys

import numpy as np
import torch

import pearl_b200
from oracle.pearl_oracle import flat
from oracle.sac_oracle import OracleSAC

obs, act, B, R = [int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (376, 17, 512, 1))]
torch.manual_seed(5)
n = 2000
rng = np.random.Generator(np.random.PCG64(3))
q8 = lambda x: (np.rint(x * 256) / 256).astype(np.float32)
low, high = -np.ones(act, dtype=np.float32), np.ones(act, dtype=np.float32)
st, ns, rw = q8(rng.standard_normal((n, obs))), q8(rng.standard_normal((n, obs))), q8(rng.standard_normal(n))
ac = q8(rng.uniform(low, high, size=(n, act)))
term = rng.random(n) < 0.05
orc = OracleSAC(obs, act, (256, 256), (256, 256), low, high, actor_lr=3e-4, critic_lr=3e-4, gamma=0.99, tau=0.005, autotune=True)
for m in [orc.actor] + orc.q:
    for mod in m.modules():
        if isinstance(mod, torch.nn.Linear):
            torch.nn.init.xavier_uniform_(mod.weight)
            mod.bias.data.fill_(0.01)
for i in range(2):
    orc.qt[i].load_state_dict(orc.q[i].state_dict())
buf = pearl_b200.B200ReplayBuffer(n)
buf.is_action_continuous = True
buf.push_batch(torch.from_numpy(st), torch.from_numpy(ac), torch.from_numpy(rw), torch.from_numpy(ns), torch.from_numpy(term),
               torch.zeros(n, dtype=torch.bool))
pl = pearl_b200.B200ContinuousSoftActorCritic(state_dim=obs, low=low, high=high, actor_hidden_dims=[256, 256], critic_hidden_dims=[256, 256],
                                              training_rounds=R, batch_size=B, actor_learning_rate=3e-4, critic_learning_rate=3e-4)
pl.load_parameters(flat(orc.actor), flat(orc.q[0]), flat(orc.q[1]))
init_actor = flat(orc.actor).clone()
noise = torch.randn(R, 2, B, act)
random.seed(77)
trace = {}
rep = pl.learn(buf, noise=noise, trace=trace)
for r in range(R):
    idx = trace["idx"][r].tolist()
    t = lambda x: torch.from_numpy(x[idx])
    out = orc.learn_batch(dict(state=t(st), action=t(ac), reward=t(rw), next_state=t(ns), terminated=t(term)), noise[r, 0], noise[r, 1])
    print("round", r, "actor_loss", rep["actor_loss"][r], out["actor_loss"], "critic", rep["critic_loss"][r], out["critic_loss"])
a, b = pl.actor_params.cpu(), flat(orc.actor)
off = 0
for name, shp in zip(("W1", "b1", "W2", "b2", "Wmu", "bmu", "Wsd", "bsd"), pl._actor_shapes()):
    k = int(np.prod(shp))
    d = (a[off:off + k] - b[off:off + k]).abs()
    upd = (b[off:off + k] - init_actor[off:off + k]).abs()
    print(f"{name:4s} n={k:6d} max_err={d.max():.3e} bad(>1.5e-5)={(d > 1.5e-5).sum().item():5d} mean|update|={upd.mean():.3e} "
          f"frac update<0.5lr={(upd < 1.5e-4 * R).float().mean():.3f}")
    if name == "W1":
        bad = (d.view(shp) > 1.5e-5)
        print("   bad per input column (top):", torch.topk(bad.sum(0).float(), 5), " per unit:", torch.topk(bad.sum(1).float(), 5))
    off += k
pc = pl.critic_params.numel() // 2
for i in range(2):
    d = (pl.critic_params[i * pc:(i + 1) * pc].cpu() - flat(orc.q[i])).abs()
    print(f"q{i + 1} max_err={d.max():.3e} bad={(d > 1.5e-5).sum().item()}")
