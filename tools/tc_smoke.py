"""Developer tool: the tensor-core learner against the SIMT learner on a small buffer (quick protocol / numerics check)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pearl_b200
from bench import Space
obs, A, B, n = (int(x) for x in (sys.argv[1:5] + [128, 16, 256, 5000][len(sys.argv) - 1:]))
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda", 0)
res = {}
for engine in ("simt", "tc"):
    g = torch.Generator(device=dev).manual_seed(1)
    b = pearl_b200.B200ReplayBuffer(n, rng="device")
    b.push_batch(torch.randn((n, obs), generator=g, device=dev), (torch.arange(n, device=dev) % A).to(torch.int32),
                 torch.randn(n, generator=g, device=dev), torch.randn((n, obs), generator=g, device=dev),
                 torch.rand(n, generator=g, device=dev) < 0.02, torch.zeros(n, dtype=torch.bool, device=dev), max_number_actions=A)
    b.seed(3)
    torch.manual_seed(7)
    L = pearl_b200.B200DeepQLearning(state_dim=obs, action_space=Space(A), hidden_dims=[64, 64], training_rounds=rounds, batch_size=B,
                                     target_update_freq=2, soft_update_tau=0.5,
                                     action_representation_module=pearl_b200.OneHotActionTensorRepresentationModule(A), engine=engine).to(dev)
    print(engine, "launching", flush=True)
    rep = L.learn(b, trace=True)
    torch.cuda.synchronize()
    res[engine] = (np.asarray(rep["loss"]), rep["q"].cpu().numpy(), rep["y"].cpu().numpy(), L.flat_parameters.cpu().numpy(),
                   L.flat_target_parameters.cpu().numpy())
    print(engine, "loss", rep["loss"], flush=True)
names = ["loss", "q", "y", "params", "target"]
for nm, x, y in zip(names, res["simt"], res["tc"]):
    err = np.max(np.abs(x - y) / (np.abs(x) + 1e-2))
    print(f"{nm:8s} max rel err {err:.3e}")
