"""Per-epoch wall clock of trpo() for both kernels (16 Ki envs x 400 steps).  Tuning tool."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml4ca_b200 as M

dev = torch.device("cuda", 0)
for kern in ("fp32", "tensor_core", "fp32", "tensor_core"):
    env = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=1 << 14, device=dev, seed=7, auto_reset=True)
    ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", device=dev, seed=7)
    marks = [time.perf_counter()]
    def mark(info):
        torch.cuda.synchronize(); marks.append(time.perf_counter())
    M.trpo(env, ac, steps_per_epoch=400, epochs=6, seed=7, graph=True, logger=mark, kernel=kern)
    print(kern, ["%.3f" % (b - a) for a, b in zip(marks[:-1], marks[1:])])
