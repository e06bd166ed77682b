"""Developer tool: phase breakdown (SM clocks of CTA 0) of the tensor-core learner kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
import pearl_b200
from pearl_b200 import _lib
from bench import Space, OBS, N_ACT, HIDDEN, BATCH
rounds = 64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
cap = 200_000
b = pearl_b200.B200ReplayBuffer(cap, rng="device")
b.push_batch(torch.randn((cap, OBS), generator=g, device=dev), (torch.arange(cap, device=dev) % N_ACT).to(torch.int32),
             torch.randn(cap, generator=g, device=dev), torch.randn((cap, OBS), generator=g, device=dev),
             torch.rand(cap, generator=g, device=dev) < 0.02, torch.zeros(cap, dtype=torch.bool, device=dev), max_number_actions=N_ACT)
b.seed(1)
L = pearl_b200.B200DeepQLearning(state_dim=OBS, action_space=Space(N_ACT), hidden_dims=list(HIDDEN), training_rounds=rounds, batch_size=BATCH,
                                 action_representation_module=pearl_b200.OneHotActionTensorRepresentationModule(N_ACT),
                                 max_rounds_per_call=rounds, engine="tc").to(dev)
L.learn(b)
st = torch.zeros((rounds, 16), dtype=torch.int64, device=dev)
_lib.check(L._libh.prl_dqn_set_profile(L._handle, C.c_void_p(st.data_ptr())))
L.set_kernel_timing(True)
L.learn(b)
ms = L.last_kernel_ms()
s = st.cpu()[4:].double()
names = ["row scalars + soft upd", "load target weights", "target layer 1 (2 tiles)", "all-actions (32 tiles)", "load online weights",
         "online layer 1", "tile0: fwd L2 + dZ2 + dH1", "rest (weight grads t0, tile1 all)", "AdamW"]
tot = (s[1:, 0] - s[:-1, 0]).mean()
print(f"kernel {ms*1e3/rounds:.1f} us/round; {tot:.0f} clk/round")
for i, n in enumerate(names):
    print(f"  {n:36s} {(s[:, i+1]-s[:, i]).mean():10.0f} clk")

print("online layer 1, group 0, chunk 1: loads+split+tmem st %.0f | fence+sync %.0f | issue %.0f | wait %.0f | (chunk0 total %.0f) | tail-to-barrier %.0f" % (
    (s[:, 11]-s[:, 10]).mean(), (s[:, 12]-s[:, 11]).mean(), (s[:, 13]-s[:, 12]).mean(), (s[:, 14]-s[:, 13]).mean(), (s[:, 10]-s[:, 5]).mean(), (s[:, 6]-s[:, 15]).mean()))
