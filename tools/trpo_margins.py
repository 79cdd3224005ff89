"""Measured distance of the TRPO update to the float64 oracle for both kernels (Hessian-vector product, CG direction, step
length, KL, line-search index) on the batches of tests/test_trpo_gpu.py.  Tuning / evidence tool."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_trpo_gpu as TT
from oracle import trpo_oracle as TO
import ml4ca_b200 as M
dev = torch.device("cuda", 0)
for seed in (2, 3, 4):
    T, n = 4, 8192
    ac, data, prob, theta, mu64 = TT._setup(dev, T, n, seed=seed)
    buf = M.GAEBuffer(9, 7, T, n, device=dev); buf.obs_buf.copy_(data[0]); buf.record_info(ac)
    full = data + [buf.log_std_buf, buf.mu_buf]
    g64, _ = prob.gradient(theta)
    ref = TO.update(prob, theta)
    for kern in ("fp32", "tensor_core"):
        upd = M.TRPOUpdater(ac, kernel=kern)
        upd._set_pi(theta)
        if kern == "tensor_core": upd._record_mu_tc(full, T, n)
        h = upd.hvp(full, T, n, theta, g64, tensor_core=(kern == "tensor_core"))
        h64 = prob.hvp(theta, g64, damping=0.1)
        info = upd.update_policy(full, T, n)
        x, x64 = upd.last["x"], ref["x"]
        cos = np.dot(x, x64) / (np.linalg.norm(x) * np.linalg.norm(x64))
        print(seed, kern, "hvp err %.4f" % (np.linalg.norm(h - h64) / np.linalg.norm(h64)), "cos %.5f" % cos,
              "alpha ratio %.4f" % (upd.last["alpha"] / ref["alpha"]), "KL %.5f (ref %.5f)" % (info["KL"], ref["kl"]),
              "bt", info.get("BacktrackIters"), ref["backtrack_iters"], "dL %.5f" % info["DeltaLossPi"])
        upd._set_pi(theta)
