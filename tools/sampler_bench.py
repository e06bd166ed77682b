import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pearl_b200
dev = torch.device("cuda", 0)
n = 1_000_000
b = pearl_b200.B200ReplayBuffer(n, rng="device")
chunk = 1 << 18
for s in range(0, n, chunk):
    m = min(chunk, n - s)
    b.push_batch(torch.zeros((m, 8), device=dev), torch.zeros(m, dtype=torch.int32, device=dev), torch.zeros(m, device=dev),
                 torch.zeros((m, 8), device=dev), torch.zeros(m, dtype=torch.bool, device=dev), torch.zeros(m, dtype=torch.bool, device=dev), max_number_actions=2)
b.seed(1)
for rounds in (256, 2048):
    b.sample_indices(256, rounds)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.sample_indices(256, rounds); e1.record(); torch.cuda.synchronize()
    print(f"rounds={rounds}: {e0.elapsed_time(e1)*1e3/rounds:.2f} us per 256-sample")
