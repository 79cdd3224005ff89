"""Learning check: PPO (or TRPO) on the batched RevoltFinal env for a few dozen epochs; prints the per-epoch reward.
Usage: [HIDDEN=80,80,80] python tools/train_demo.py [ppo|trpo] [epochs] [num_envs] [graph|eager] [fp32|tensor_core]
(HIDDEN: hidden layer sizes; default 64,64 -- 80,80,80 is the reference's own network, train.py:30-32)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml4ca_b200 as M

algo = sys.argv[1] if len(sys.argv) > 1 else "ppo"
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
graph = (sys.argv[4] == "graph") if len(sys.argv) > 4 else False
kw = dict(kernel=sys.argv[5]) if len(sys.argv) > 5 and algo == "trpo" else {}
dev = torch.device("cuda", 0)
env = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=n, device=dev, seed=0, auto_reset=True)
t0 = time.time()
fn = M.ppo if algo == "ppo" else M.trpo
hidden = tuple(int(x) for x in os.environ.get("HIDDEN", "64,64").split(","))
ac, hist = fn(env, steps_per_epoch=400, epochs=epochs, seed=0, graph=graph, hidden_sizes=hidden, **kw)
torch.cuda.synchronize()
for h in hist:
    if h["Epoch"] % max(1, epochs // 20) == 0 or h["Epoch"] == epochs - 1:
        print("epoch %3d  step reward %8.4f  EpRet %9.2f  EpLen %6.1f  VVals %8.2f  KL %.4f  LossV %9.2f" % (
            h["Epoch"], h["AverageStepReward"], h["AverageEpRet"], h["EpLen"], h["AverageVVals"], h["KL"], h["LossV"]))
print("%s%s hidden %s: %d epochs x %d envs x 400 steps in %.1f s" % (algo, " (graph rollout)" if graph else "", hidden, epochs, n, time.time() - t0))
