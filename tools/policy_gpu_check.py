import sys, numpy as np, torch
sys.path.insert(0, '.')
import ml4ca_b200 as M
from oracle import mlp_oracle as MO
torch.manual_seed(0)
for tag, dims, act in [('64x64', dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2), 'leaky_relu'),
                       ('final_80x3', None, 'leaky_relu'), ('limited_64x3', None, 'leaky_relu'),
                       ('64x64 tanh', dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2), 'tanh')]:
    if dims is None:
        g = np.load('tests/golden/policy_%s.npz' % tag)
        dims = {k: int(g[k]) for k in ('obs_dim', 'act_dim', 'hidden', 'n_hidden')}
        flat = g['params']
    else:
        flat = MO.glorot_params(dims, seed=3)
    ac = M.ActorCritic(dims['obs_dim'], dims['act_dim'], (dims['hidden'],) * dims['n_hidden'], act, params=flat)
    for n in (1, 100, 128 * 5 + 17, 100000):
        obs = (torch.rand(dims['obs_dim'], n, device='cuda') * 2 - 1) * torch.tensor([8, 8, .8, 1.4, .3, .5, 1, 1, 1.], device='cuda')[:dims['obs_dim'], None]
        pi, v, logp, mu = ac.step(obs, deterministic=True, return_mu=True)
        torch.cuda.synchronize()
        ref = MO.forward(flat, dims, obs.cpu().numpy(), act)
        emu = np.abs(mu.cpu().numpy() - ref['mu']).max(); ev = np.abs(v.cpu().numpy() - ref['v']).max()
        print(tag, 'n', n, 'max|mu err|', emu, 'mu scale', np.abs(ref['mu']).max(), 'max|v err|', ev, 'v scale', np.abs(ref['v']).max())
# throughput
dims = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)
ac = M.ActorCritic(9, 7, (64, 64), 'leaky_relu', params=MO.glorot_params(dims, 3))
n = 1 << 23
obs = torch.rand(9, n, device='cuda') * 2 - 1
out = (torch.empty(7, n, device='cuda'), torch.empty(n, device='cuda'), torch.empty(n, device='cuda'))
for _ in range(3): ac.step(obs, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ac.step(obs, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('policy forward 8Mi obs: %.3f ms -> %.2f G obs/s, %.1f TFLOP/s' % (ms, n / ms / 1e6, n * 19712 / ms / 1e9))
