"""Stage timeline of the actor/critic forward kernel on SM 0 (needs a library built with -DML4CA_POLICY_TRACE, passed
through ML4CA_LIB).  Prints, per tile group and round, the clock64 stamps: 0 tile start, 1 operands staged, 2 layer-1
accumulators ready, 3 epilogue 1 done, 4 layer-2 ready, 5 epilogue 2 done, 6 output ready, 7 tile done.  Tuning tool."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ml4ca_b200 as M


n = 1 << 23
dev = torch.device("cuda", 0)
ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", device=dev, seed=3)
obs = torch.rand(9, n, device=dev) * 2 - 1
out = (torch.empty(7, n, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev))
for i in range(3):
    ac.step(obs, out=out, step=i)
torch.cuda.synchronize()
L = ctypes.CDLL(os.environ["ML4CA_LIB"])
buf = (ctypes.c_longlong * (4 * 64 * 16))()
assert L.ml4ca_debug_policy_trace(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(4, 64, 16)
t0 = t[:, 0, 0].min()
for r in list(range(0, 4)) + list(range(40, 46)):
    for g in range(4):
        rel = t[g, r] - t0
        print("round %2d group %d  start %7d | stage %5d | L1 wait %5d | epi1 %5d | L2 wait %5d | epi2 %5d | out wait %5d | epi3 %5d | total %5d"
              % (r, g, rel[0], rel[1] - rel[0], rel[2] - rel[1], rel[3] - rel[2], rel[4] - rel[3], rel[5] - rel[4], rel[6] - rel[5],
                 rel[7] - rel[6], rel[7] - rel[0]))
d = np.diff(t[:, :, :8], axis=2)[:, 8:60]
print("mean per stage (rounds 8..59):", np.round(d.mean(axis=(0, 1))).astype(int), "tile total", int((t[:, 8:60, 7] - t[:, 8:60, 0]).mean()))
m = lambda a, b: int((t[:, 8:60, a] - t[:, 8:60, b]).mean())
print("hand-over 1 (fence + barrier + issue): %d, then wait %d | hand-over 2: %d, noise+logp %d, then wait %d"
      % (m(8, 3), m(4, 8), m(9, 5), m(10, 9), m(6, 10)))
print("hand-over 2 split: proxy fence %d | group barrier %d | elect + issue %d" % (m(11, 5), m(12, 11), m(9, 12)))
print("  inside: barrier -> elected %d | fence + MMAs %d | commit %d | elected -> reconverged %d" % (m(13, 12), m(14, 13), m(15, 14), m(9, 15)))
print("period per group (start to next start):", np.round(np.diff(t[:, 8:60, 0], axis=1).mean(axis=1)).astype(int))
