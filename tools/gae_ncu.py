import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pearl_b200.ppo import gae_and_lambda_returns
dev = torch.device("cuda", 0)
ng = 1 << 24
g = torch.Generator(device=dev).manual_seed(0)
vals, rws = torch.randn(ng, device=dev, generator=g), torch.randn(ng, device=dev, generator=g)
te = ((torch.arange(ng, device=dev) % 500) == 499).to(torch.uint8)
tu = torch.zeros(ng, dtype=torch.uint8, device=dev)
for _ in range(3):
    gae_and_lambda_returns(vals, 0.1, rws, te, tu, 0.99, 0.95)
torch.cuda.synchronize()
