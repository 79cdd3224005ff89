"""Driver for the K1 ncu capture: one QP allocator launch on n demands of the config-1 law (n from argv, default 64 Ki)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ml4ca_b200 as M
from ml4ca_b200 import synth

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
dev = torch.device("cuda", 0)
tau, prev = synth.qp_batch(4096, seed=1)
reps = m // 4096
t_big = torch.as_tensor(np.tile(np.asarray(tau), reps), dtype=torch.float32, device=dev).contiguous()
tb = M.QPTA(num_envs=m, device=dev)
for i in range(2):
    tb.previous_thruster_state = np.tile(np.asarray(prev), reps)
    x, ok = tb.solve_QP(t_big)
torch.cuda.synchronize()
print("ok %.4f" % float(ok.float().mean()))
