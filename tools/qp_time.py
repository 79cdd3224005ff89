"""CUDA-event timing of the QP allocator kernel alone: 1 Mi allocations of the config-1 demand law.  Tuning tool
(ML4CA_LIB selects a variant library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ml4ca_b200 as M
from ml4ca_b200 import synth

m = 1 << 20
dev = torch.device("cuda", 0)
tau, prev = synth.qp_batch(4096, seed=1)
reps = m // 4096
t_big = torch.as_tensor(np.tile(np.asarray(tau), reps), dtype=torch.float32, device=dev).contiguous()
prev_big = np.tile(np.asarray(prev), reps)
tb = M.QPTA(num_envs=m, device=dev)
ok_rate = None
times = []
for i in range(6):
    tb.previous_thruster_state = prev_big
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    x, ok = tb.solve_QP(t_big)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
    ok_rate = float(ok.float().mean())
print(os.path.basename(os.environ.get("ML4CA_LIB", "libml4ca_b200.so")), os.environ.get("ML4CA_QP_THREADS", "-"),
      "ms %.3f (min of %s)" % (min(times[1:]), ["%.2f" % t for t in times]), "M alloc/s %.1f" % (m / min(times[1:]) / 1e3),
      "success %.4f" % ok_rate, "checksum %.6f" % float(x.double().abs().mean()),
      "sha %s" % __import__("hashlib").sha256(x.cpu().numpy().tobytes()).hexdigest()[:12])
