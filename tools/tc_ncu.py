"""Developer tool: a short single-learner run of the tensor-core learner for `ncu --set full -k regex:k_dqn_tc`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pearl_b200
from bench import Space, OBS, N_ACT, HIDDEN, BATCH
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cap = 100_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
bufs, ls = [], []
for i in range(R):
    b = pearl_b200.B200ReplayBuffer(cap, rng="device")
    b.push_batch(torch.randn((cap, OBS), generator=g, device=dev), (torch.arange(cap, device=dev) % N_ACT).to(torch.int32),
                 torch.randn(cap, generator=g, device=dev), torch.randn((cap, OBS), generator=g, device=dev),
                 torch.rand(cap, generator=g, device=dev) < 0.02, torch.zeros(cap, dtype=torch.bool, device=dev), max_number_actions=N_ACT)
    b.seed(i + 1)
    bufs.append(b)
    ls.append(pearl_b200.B200DeepQLearning(state_dim=OBS, action_space=Space(N_ACT), hidden_dims=list(HIDDEN), training_rounds=rounds, batch_size=BATCH,
                                           action_representation_module=pearl_b200.OneHotActionTensorRepresentationModule(N_ACT),
                                           max_rounds_per_call=rounds, engine="tc").to(dev))
grp = pearl_b200.B200LearnerGroup(ls, bufs)
for _ in range(3):
    grp.learn()
torch.cuda.synchronize()
print("done")
