"""Wall-clock of one 400-step rollout (policy forward + env step per step), eager launches against CUDA-graph replay, for a
few batch sizes.  Tuning tool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml4ca_b200 as M

dev = torch.device("cuda", 0)
ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", device=dev, seed=1)
T = 400
for n in (4, 1024, 16384, 262144):
    res = {}
    for graph in (False, True):
        env = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=n, device=dev, seed=0, auto_reset=True)
        env.reset()
        buf = M.TrajectoryBuffer(9, 7, T, n, device=dev)
        for i in range(3):
            M.rollout(env, ac, buf, seed=0, start_step=i * T, graph=graph)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(5):
            M.rollout(env, ac, buf, seed=0, start_step=(3 + i) * T, graph=graph)
        torch.cuda.synchronize()
        res[graph] = (time.perf_counter() - t0) / 5
    print("n %7d  eager %.2f ms  graph %.2f ms per 400-step rollout  (%.1f / %.1f us per step)" % (
        n, res[False] * 1e3, res[True] * 1e3, res[False] / T * 1e6, res[True] / T * 1e6))
