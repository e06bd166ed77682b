"""Developer tool: aggregate rate of the tensor-core learner group (R learners, `rounds` rounds) + phase stamps of CTA 0."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
import pearl_b200
from pearl_b200 import _lib
from bench import Space, OBS, N_ACT, HIDDEN, BATCH
R = int(sys.argv[1]) if len(sys.argv) > 1 else 144
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
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
grp.set_kernel_timing(True)
for _ in range(2):
    grp.learn()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    grp.learn()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"R={R} rounds={rounds}: {R*rounds/dt:.4e} steps/s wall; kernel {grp.last_kernel_ms():.2f} ms -> {R*rounds/grp.last_kernel_ms()*1e3:.4e} steps/s; {grp.last_kernel_ms()*1e3/rounds:.1f} us/round")
# phase stamps (single learner launch, unchunked)
st = torch.zeros((rounds, 16), dtype=torch.int64, device=dev)
_lib.check(ls[0]._libh.prl_dqn_set_profile(ls[0]._handle, C.c_void_p(st.data_ptr())))
grp.learn()
torch.cuda.synchronize()
s = st.cpu()[4:].double()
names = ["soft upd + barrier A", "target layer 1", "all-actions", "online layer 1 (+h1)", "L2 + loss + dH1", "barrier B + zero", "weight-grad passes", "AdamW"]
tot = (s[1:, 0] - s[:-1, 0]).mean()
print(f"{tot:.0f} clk/round (CTA 0, group 0)")
for i, n in enumerate(names):
    print(f"  {n:28s} {(s[:, i+1]-s[:, i]).mean():10.0f} clk")
d = lambda a, b: (s[:, a] - s[:, b]).mean()
print(f"  AdamW region: next row + chunks {d(9,7):.0f} | G1 part {d(10,9):.0f} | G2 / W3 part {d(11,10):.0f} | rows_sync {d(12,11):.0f} | smalls {d(8,12):.0f}")
print(f"  action slot 2: epilogue(0)+loop {d(13,2):.0f} since phase start | wait MMA {d(14,13):.0f} | build + publish {d(15,14):.0f}")
