import sys
sys.path.insert(0, '.')
import numpy as np, torch
import ml4ca_b200 as M
taus = np.array([[0, -2, -2], [0, -2, 2], [-1, -2, 0.5], [0.0, -1.5, 0.0], [4, 0, 0], [0.0, -0.3, 0.1]], dtype=np.float32).T
n = taus.shape[1]
qp = M.QPTA(num_envs=n)
x, ok = qp.solve_QP(torch.as_tensor(taus, device='cuda'))
print(ok.cpu().numpy(), (qp.last_status.cpu().numpy().astype(np.uint32) >> 24))
print(np.round(x.cpu().numpy().T, 3))
out = qp.tau_controller_callback_func(torch.as_tensor(taus, device='cuda'))
print(np.round(qp.last_output.cpu().numpy().T, 3))
print(np.round(qp._prev.cpu().numpy().T, 3))
