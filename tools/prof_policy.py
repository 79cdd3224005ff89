import sys, numpy as np, torch
sys.path.insert(0, '.')
import ml4ca_b200 as M
from oracle import mlp_oracle as MO
dims = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)
ac = M.ActorCritic(9, 7, (64, 64), 'leaky_relu', params=MO.glorot_params(dims, 3))
n = 1 << 22
obs = torch.rand(9, n, device='cuda') * 2 - 1
out = (torch.empty(7, n, device='cuda'), torch.empty(n, device='cuda'), torch.empty(n, device='cuda'))
for _ in range(4): ac.step(obs, out=out)
torch.cuda.synchronize()
print('ok')
