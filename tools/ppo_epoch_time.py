"""Epoch time of ppo() -- rollout + GAE + update -- with the host-driven update loop and with the update replayed from one
CUDA graph (device-side KL stop), at the bench size (16 Ki envs x 400 steps) and at the reference's own batch (4 x 400).
    python tools/ppo_epoch_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml4ca_b200 as M

dev = torch.device("cuda", 0)
for hidden in ((64, 64), (80, 80, 80)):
    for ne in (1 << 14, 4):
        for upd_graph in (False, True):
            env = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=ne, device=dev, seed=5, auto_reset=True)
            marks = []

            def mark(info):
                torch.cuda.synchronize()
                marks.append(time.perf_counter())
            epochs = 6 if ne > 100 or hidden == (64, 64) else 5
            if hidden == (80, 80, 80) and ne > 100:
                epochs = 4
            _, hist = M.ppo(env, steps_per_epoch=400, epochs=epochs, seed=5, graph=True, update_graph=upd_graph, hidden_sizes=hidden,
                            logger=mark)
            dts = sorted(b - a for a, b in zip(marks[1:], marks[2:])) or [float("nan")]
            print("hidden %s  %6d envs x 400: update %-6s epoch %.2f ms (median of %d)  StopIter %s" % (
                hidden, ne, "graph" if upd_graph else "eager", 1e3 * dts[len(dts) // 2], len(dts), [h["StopIter"] for h in hist]))
