"""Development check: tensor-core PPO gradient kernel against the fp32 CUDA-core kernel on the same batch."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import ml4ca_b200 as M
from ml4ca_b200 import _lib
dev = torch.device('cuda', 0)
T, n = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 1000
act = sys.argv[3] if len(sys.argv) > 3 else 'leaky_relu'
g = torch.Generator(device=dev); g.manual_seed(0)
ac = M.ActorCritic(9, 7, (64, 64), act, device=dev, seed=1)
p = ac.parameters(); p.add_(torch.randn(p.shape, device=dev, generator=g) * 0.05); ac.refresh()
data = (torch.randn(T, 9, n, device=dev, generator=g), torch.randn(T, 7, n, device=dev, generator=g),
        torch.randn(T, n, device=dev, generator=g), torch.randn(T, n, device=dev, generator=g) * 3,
        torch.randn(T, n, device=dev, generator=g) * 0.5 - 9.0)
upd = M.PPOUpdater(ac)
res = {}
for fp32 in (1, 0):
    _lib.lib().ml4ca_ppo_use_fp32(fp32)
    for net in (0, 1):
        s, c = upd._grad(net, data, T, n)
        torch.cuda.synchronize()
        res[(fp32, net)] = (upd.flat[:ac.num_params].clone(), s)
npi = ac.var_counts[0]
for net, sl in ((0, slice(0, npi)), (1, slice(npi, None))):
    a, b = res[(1, net)][0][sl], res[(0, net)][0][sl]
    print('net', net, 'max|g32|', a.abs().max().item(), 'max diff', (a - b).abs().max().item(), 'rel', ((a - b).abs().max() / a.abs().max()).item())
    print('   stats fp32', [round(x, 4) for x in res[(1, net)][1]], '\n   stats tc  ', [round(x, 4) for x in res[(0, net)][1]])
    if True:
        names = [('w1', 0, 576), ('b1', 576, 640), ('w2', 640, 4736), ('b2', 4736, 4800), ('wo', 4800, 5248), ('bo', 5248, 5255), ('ls', 5255, 5262)] if net == 0 else \
                [('w1', 0, 576), ('b1', 576, 640), ('w2', 640, 4736), ('b2', 4736, 4800), ('wo', 4800, 4864), ('bo', 4864, 4865)]
        for nm, lo, hi in names:
            print('   %s: max|a| %.4g maxdiff %.4g' % (nm, a[lo:hi].abs().max().item(), (a[lo:hi] - b[lo:hi]).abs().max().item()))
