"""Iteration histogram of the QP kernel on the config-1 demand law (status word bits 24-31 = SQP iterations), split by
outcome.  Tuning tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ml4ca_b200 as M
from ml4ca_b200 import synth

n = 1 << 16
tau, prev = synth.qp_batch(n, seed=1)
ta = M.QPTA(num_envs=n)
ta.previous_thruster_state = np.asarray(prev)
x, ok = ta.solve_QP(torch.as_tensor(np.asarray(tau), dtype=torch.float32, device="cuda"))
st = ta.last_status.cpu().numpy().astype(np.uint32)
ok = ok.cpu().numpy()
its = (st >> 24).astype(np.int64)
print("success rate %.4f" % ok.mean())
print("iterations, successes:", np.bincount(its[ok], minlength=27)[:27])
print("iterations, failures :", np.bincount(its[~ok], minlength=27)[:27])
print("mean iterations: success %.2f, failure %.2f; share of all iterations spent on failures %.2f"
      % (its[ok].mean(), its[~ok].mean(), its[~ok].sum() / its.sum()))
