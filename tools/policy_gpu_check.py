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

# ---- fused rollout step vs (policy forward -> env step) as separate kernels -------------------------------------
import ctypes
from ml4ca_b200 import _lib
from ml4ca_b200.env import RevoltFinal, StandInHull
n = 100000
envA = RevoltFinal(StandInHull(), extended_state=True, cont_ang=True, num_envs=n, seed=5, auto_reset=True, max_ep_len=60)
envB = RevoltFinal(StandInHull(), extended_state=True, cont_ang=True, num_envs=n, seed=5, auto_reset=True, max_ep_len=60)
oA = envA.reset(); oB = envB.reset()
ac.seed = 99
dev = torch.device('cuda')
T = 8
for t in range(T):
    # separate
    pi, v, logp = ac.step(oA, step=t)
    oA2, rA, dA, infoA = envA.step(pi)
    # fused
    obs_f = torch.empty(9, n, device=dev); act_f = torch.empty(7, n, device=dev); rew_f = torch.empty(n, device=dev)
    val_f = torch.empty(n, device=dev); logp_f = torch.empty(n, device=dev); done_f = torch.empty(n, dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().ml4ca_rollout_step(envB._handle, ac._handle, 99, t, 0, _lib.ptr(obs_f), _lib.ptr(act_f), _lib.ptr(rew_f),
                                             _lib.ptr(val_f), _lib.ptr(logp_f), _lib.ptr(done_f), _lib.current_stream()))
    torch.cuda.synchronize()
    print('t', t, 'obs', float((obs_f - oA).abs().max()), 'act', float((act_f - pi).abs().max()), 'val', float((val_f - v).abs().max()),
          'logp', float((logp_f - logp).abs().max()), 'rew', float((rew_f - rA).abs().max()), 'done eq', bool((done_f == infoA['flags']).all()),
          'n done', int((done_f != 0).sum()))
    oA = oA2
sA, sB = envA.get_state(), envB.get_state()
print('state diff', {k: float((sA[k].float() - sB[k].float()).abs().max()) for k in sA})
# fused throughput, 16 Mi envs, training-mode records
n = 1 << 24
env = RevoltFinal(StandInHull(), extended_state=True, cont_ang=True, num_envs=n, seed=2, auto_reset=True)
env.reset()
bufs = [torch.empty(9, n, device=dev), torch.empty(7, n, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)]
def fused(t, rec=True):
    ptrs = [_lib.ptr(b) if rec else None for b in bufs]
    if not rec: ptrs[2] = _lib.ptr(bufs[2]); ptrs[5] = _lib.ptr(bufs[5])
    _lib.check(_lib.lib().ml4ca_rollout_step(env._handle, ac._handle, 7, t, 0, *ptrs, _lib.current_stream()))
for rec in (True, False):
    for t in range(3): fused(t, rec)
    e0.record()
    for t in range(20): fused(t, rec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('fused rollout step 16Mi envs (%s): %.3f ms -> %.2f G env-steps/s' % ('training records' if rec else 'inference', ms, n / ms / 1e6))
