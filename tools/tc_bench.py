"""Developer tool: throughput of the tensor-core learner group for several group sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pearl_b200
from bench import Space, OBS, N_ACT, HIDDEN, BATCH

cap = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sizes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 8, 37, 74, 147]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
maxL = max(sizes)
bufs, learners = [], []
for i in range(maxL):
    b = pearl_b200.B200ReplayBuffer(cap, rng="device")
    b.push_batch(torch.randn((cap, OBS), generator=g, device=dev), (torch.arange(cap, device=dev) % N_ACT).to(torch.int32),
                 torch.randn(cap, generator=g, device=dev), torch.randn((cap, OBS), generator=g, device=dev),
                 torch.rand(cap, generator=g, device=dev) < 0.02, torch.zeros(cap, dtype=torch.bool, device=dev),
                 max_number_actions=N_ACT)
    b.seed(i)
    bufs.append(b)
    learners.append(pearl_b200.B200DeepQLearning(
        state_dim=OBS, action_space=Space(N_ACT), hidden_dims=list(HIDDEN), training_rounds=rounds, batch_size=BATCH,
        action_representation_module=pearl_b200.OneHotActionTensorRepresentationModule(N_ACT), max_rounds_per_call=rounds,
        engine="tc").to(dev))
for L in sizes:
    grp = pearl_b200.B200LearnerGroup(learners[:L], bufs[:L])
    grp.set_kernel_timing(True)
    grp.learn()
    t0 = time.perf_counter(); grp.learn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ms = grp.last_kernel_ms()
    print(f"L={L:4d}: kernel {ms*1e3/rounds:8.2f} us/round  -> {L*rounds/(ms/1e3):12.0f} steps/s (kernel)  {L*rounds/dt:12.0f} steps/s (wall incl. sampler)")
