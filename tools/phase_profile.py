"""Developer tool: per-phase SM-clock breakdown of the persistent learner kernel
(CTA 0), printed in microseconds at the measured SM clock."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
import pearl_b200
from pearl_b200 import _lib
from bench import Space, OBS, N_ACT, HIDDEN, BATCH

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda", 0)
buf = pearl_b200.B200ReplayBuffer(200_000, rng="device")
g = torch.Generator(device=dev).manual_seed(1)
m = 200_000
buf.push_batch(torch.randn((m, OBS), generator=g, device=dev), (torch.arange(m, device=dev) % N_ACT).to(torch.int32),
               torch.randn(m, generator=g, device=dev), torch.randn((m, OBS), generator=g, device=dev),
               torch.rand(m, generator=g, device=dev) < 0.02, torch.zeros(m, dtype=torch.bool, device=dev),
               max_number_actions=N_ACT)
buf.seed(1)
L = pearl_b200.B200DeepQLearning(state_dim=OBS, action_space=Space(N_ACT), hidden_dims=list(HIDDEN), training_rounds=rounds,
                                 batch_size=BATCH, action_representation_module=pearl_b200.OneHotActionTensorRepresentationModule(N_ACT),
                                 max_rounds_per_call=rounds, rows_per_cta=rows).to(dev)
L.learn(buf)
stamps = torch.zeros((rounds, 16), dtype=torch.int64, device=dev)
_lib.check(L._libh.prl_dqn_set_profile(L._handle, C.c_void_p(stamps.data_ptr())))
L.set_kernel_timing(True)
L.learn(buf)
ms = L.last_kernel_ms()
s = stamps.cpu()[8:].double()
names = ["wait records", "scalars+action cols", "online L1 (+h1)", "online L2 + head", "target L1", "all-actions Q", "bellman", "backward dZ", "outer products + grads",
         "grid barrier 1", "phase B update", "grid barrier 2"]
d = (s[:, 1:13] - s[:, 0:12]).mean(0)
tot = (s[1:, 0] - s[:-1, 0]).mean()
mhz = tot / (ms * 1e3 / rounds)
print(f"kernel {ms*1e3/rounds:.2f} us/round, {tot:.0f} clk/round -> {mhz:.0f} MHz; info {L.launch_info()}")
for n, v in zip(names, d):
    print(f"  {n:28s} {v:9.0f} clk  {v/mhz:7.2f} us")
sub = [("  .. prefetch issue", s[:, 13] - s[:, 8]), ("  .. cta_outer x2", s[:, 14] - s[:, 13]), ("  .. dW1 action cols", s[:, 15] - s[:, 14]),
       ("  .. biases/w3/mae", s[:, 9] - s[:, 15])]
for n, v in sub:
    print(f"  {n:28s} {v.mean():9.0f} clk")
