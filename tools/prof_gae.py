"""Profiling driver for K5 (GAE-lambda) and K2 (pseudoinverse + PID): a few launches each at the bench sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml4ca_b200 as M
from ml4ca_b200 import synth

dev = torch.device("cuda", 0)
T, n = 64, 1 << 21
buf = M.TrajectoryBuffer(1, 1, T, n, gamma=0.99, lam=0.97, device=dev)
buf.rew_buf.normal_(); buf.val_buf.normal_()
for _ in range(4):
    buf.finish_path()
m = 1 << 20
eta, nu, ref, _ = synth.pose_batch(m, seed=1)
t = lambda x: torch.as_tensor(x, dtype=torch.float32, device=dev).contiguous()
integ = torch.zeros(3, m, device=dev)
for _ in range(4):
    M.pinv_pid(t(eta), t(nu), t(ref), integ)
torch.cuda.synchronize()
print("ok")
