"""Development check of the CUDA allocator against the float64 prototype and the reference restatement."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import ml4ca_b200 as M
from oracle import qp_oracle as QO
from tools import qp_sqp_proto as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
tau, prev = QO.synth_batch(n, seed=0)
tau, prev = tau.astype(np.float32).astype(np.float64), prev.astype(np.float32).astype(np.float64)
ta = M.QPTA(num_envs=n)
ta.previous_thruster_state = prev
x, ok = ta.solve_QP(torch.as_tensor(tau, dtype=torch.float32, device='cuda'))
torch.cuda.synchronize()
x = x.cpu().numpy().astype(np.float64); ok = ok.cpu().numpy(); st = ta.last_status.cpu().numpy().astype(np.uint32)
its = st >> 24
print('gpu success rate', ok.mean(), 'iters hist', np.bincount(its)[:27])
d = []; bad = 0; kk = []
for j in range(n):
    xp, okp, nit, mu = P.solve(tau[:, j], prev[:, j])
    if okp != ok[j]:
        bad += 1
        if bad < 8: print('success mismatch vs proto', j, ok[j], okp, 'gpu its', its[j], 'proto its', nit)
        continue
    if okp:
        xpc = xp.copy(); xpc[np.abs(xpc) < 0.01] = 0
        e = np.abs(x[:, j] - xpc) / np.maximum(1, np.abs(xpc)); d.append(e.max())
        if e.max() > 1e-4 and len(kk) < 8: kk.append((j, e.max(), its[j], nit))
d = np.array(d)
print('vs float64 prototype: mismatched success', bad, 'median', np.median(d), 'p90', np.percentile(d, 90), 'p99', np.percentile(d, 99), 'max', d.max(), 'n>1e-5', (d > 1e-5).sum(), 'n>5e-3', (d > 5e-3).sum())
print('large diffs', kk)
# throughput
m = 1 << 20
tau2, prev2 = QO.synth_batch(4096, seed=1)
reps = m // 4096
t_big = torch.as_tensor(np.tile(tau2, reps), dtype=torch.float32, device='cuda').contiguous()
tb = M.QPTA(num_envs=m); tb.previous_thruster_state = np.tile(prev2, reps)
for _ in range(2): tb.solve_QP(t_big)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): tb.solve_QP(t_big)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('1Mi allocations: %.3f ms  -> %.1f M alloc/s' % (ms, m / ms / 1e3))
