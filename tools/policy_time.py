"""CUDA-event timing of the actor/critic forward kernel (K4) alone: 8 Mi observations, 64 x 64 leaky-ReLU nets.
Tuning tool (ML4CA_LIB selects a variant library), not a bench."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml4ca_b200 as M


n = int(os.environ.get("N", 1 << 23))
dev = torch.device("cuda", 0)
ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", device=dev, seed=3)
obs = torch.rand(9, n, device=dev) * 2 - 1
out = (torch.empty(7, n, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev))
for i in range(5):
    ac.step(obs, out=out, step=i)
torch.cuda.synchronize()
reps = 40
det = bool(int(os.environ.get("DET", "0")))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    ac.step(obs, out=out, step=10 + i, deterministic=det)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(os.path.basename(os.environ.get("ML4CA_LIB", "libml4ca_b200.so")), "det", det, "n", n, "ms %.4f" % ms, "G obs/s %.2f" % (n / ms / 1e6),
      "checksum %.6f" % float(out[0].double().mean()))
