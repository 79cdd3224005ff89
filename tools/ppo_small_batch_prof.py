"""Where one PPO epoch of the reference's own batch (4 envs x 400 steps, ppo.py defaults) spends its time: wall clock around each
phase of run_epochs (a device synchronize on both sides), graph replays for rollout and update.  GPU box tool.
    python tools/ppo_small_batch_prof.py [n_envs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml4ca_b200 as M
from ml4ca_b200 import _lib

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
for hidden in ((64, 64), (80, 80, 80)):
    env = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=ne, device=dev, seed=5, auto_reset=True)
    ac = M.ActorCritic(9, 7, hidden, "leaky_relu", device=dev, seed=5)
    buf = M.TrajectoryBuffer(9, 7, 400, ne, 0.99, 0.97, device=dev, max_ep_len=env.max_ep_len)
    upd = M.PPOUpdater(ac, target_kl=1.0)            # never stops early: all 80 + 80 iterations
    env.reset()
    t = {}

    def timed(name, fn, reps=1):
        torch.cuda.synchronize()
        a = time.perf_counter()
        for _ in range(reps):
            r = fn()
        torch.cuda.synchronize()
        t.setdefault(name, []).append((time.perf_counter() - a) / reps * 1e3)
        return r

    for epoch in range(6):
        o = timed("rollout (graph)", lambda: M.rollout(env, ac, buf, seed=5, start_step=400 * epoch, graph=True))
        v = timed("bootstrap forward", lambda: ac.step(o, deterministic=True, step=0)[1])
        timed("finish_path (GAE)", lambda: buf.finish_path(last_val=v))
        timed("env.reset", lambda: env.reset())
        l0 = _lib.launch_count()
        timed("update (graph)", lambda: upd.update(buf, graph=True))
        if epoch == 5:
            timed("update (eager)", lambda: upd.update(buf, graph=False))
    print("hidden %s, %d envs x 400 steps" % (hidden, ne))
    for k, v in t.items():
        v = sorted(v[1:] if len(v) > 2 else v)
        print("  %-20s %8.3f ms" % (k, v[len(v) // 2]))
