"""Driver of the K6 ncu captures: three gradient passes of the pi-net over 1 Mi samples (HIDDEN=80,80,80: the reference's default network; ML4CA_PPO_FP32=1 selects the fp32 kernels)."""
import os, sys
sys.path.insert(0, '.')
import torch
import ml4ca_b200 as M
dev = torch.device('cuda', 0)
T, n = 16, 1 << 16
hidden = tuple(int(x) for x in os.environ.get('HIDDEN', '64,64').split(','))
ac = M.ActorCritic(9, 7, hidden, 'leaky_relu', device=dev, seed=1)
g = torch.Generator(device=dev); g.manual_seed(0)
data = (torch.randn(T, 9, n, device=dev, generator=g), torch.randn(T, 7, n, device=dev, generator=g),
        torch.randn(T, n, device=dev, generator=g), torch.randn(T, n, device=dev, generator=g),
        torch.randn(T, n, device=dev, generator=g) - 9.0)
upd = M.PPOUpdater(ac)
for _ in range(3):
    upd._grad(0, data, T, n)
torch.cuda.synchronize()
print('ok')
