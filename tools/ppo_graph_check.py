"""How far apart do two PPO updates of the same batch end up?  (GPU box tool.)

The gradient kernels add per-CTA partial sums with atomicAdd, so two runs of the same update differ in the last bits of every
gradient; Adam then amplifies a component whose gradient is at noise level.  This prints the parameter distance after 3 epochs of
(12 pi + 9 v) iterations for eager vs eager, graph vs graph and eager vs graph, so that the test tolerance of
tests/test_ppo_update_gpu.py::test_graph_update_equals_eager_update rests on a measured noise floor.

    python tools/ppo_graph_check.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ml4ca_b200 as M  # noqa: E402


def batch(T, n, seed):
    rng = np.random.default_rng(seed)
    scale = np.array([2, 2, .3, .5, .1, .2, .5, .5, .5], dtype=np.float32)[None, :, None]
    obs = rng.normal(size=(T, 9, n)).astype(np.float32) * scale
    act = rng.normal(size=(T, 7, n)).astype(np.float32)
    adv = rng.normal(size=(T, n)).astype(np.float32)
    ret = (rng.normal(size=(T, n)) * 3).astype(np.float32)
    return obs, act, adv, ret


def run(hidden, target_kl, use_graph, dev, epochs=3):
    T, n = 4, 4096
    obs, act, adv, ret = batch(T, n, 17)
    ac = M.ActorCritic(9, 7, hidden, "leaky_relu", device=dev, seed=4)
    p0 = ac.parameters().clone()
    buf = M.TrajectoryBuffer(9, 7, T, n, device=dev)
    upd = M.PPOUpdater(ac, train_pi_iters=12, train_v_iters=9, target_kl=target_kl)
    stops = []
    for epoch in range(epochs):
        buf.obs_buf.copy_(torch.as_tensor(obs)); buf.act_buf.copy_(torch.as_tensor(act))
        buf.adv_buf.copy_(torch.as_tensor(adv) * (1.0 + 0.1 * epoch)); buf.ret_buf.copy_(torch.as_tensor(ret))
        # actions sampled from the current policy with their own log-likelihood: the ratio starts at 1
        o = buf.obs_buf.permute(1, 0, 2).reshape(9, -1).contiguous()
        a, _, lp = ac.step(o, deterministic=False, step=epoch)
        buf.act_buf.copy_(a.reshape(7, T, n).permute(1, 0, 2))
        buf.logp_buf.copy_(lp.reshape(T, n))
        stops.append(upd.update(buf, graph=use_graph)["StopIter"])
    return ac.parameters().clone(), p0, stops


def main():
    dev = torch.device("cuda", 0)
    for hidden in ((64, 64), (80, 80, 80)):
        for kl in (1.0,):
            runs = {}
            for name, g in (("eager1", False), ("eager2", False), ("graph1", True), ("graph2", True)):
                runs[name] = run(hidden, kl, g, dev)
            p0 = runs["eager1"][1]
            moved = (runs["eager1"][0] - p0).norm().item()
            print("hidden", hidden, "target_kl", kl, "stops", {k: v[2] for k, v in runs.items()}, "|moved| %.4g" % moved)
            for a, b in (("eager1", "eager2"), ("graph1", "graph2"), ("eager1", "graph1"), ("eager2", "graph2")):
                d = runs[a][0] - runs[b][0]
                print("  %s vs %s: rel L2 %.3e  max %.3e  frac<2e-5 %.4f  frac==0 %.4f" % (
                    a, b, d.norm().item() / moved, d.abs().max().item(), (d.abs() < 2e-5).float().mean().item(),
                    (d == 0).float().mean().item()))


if __name__ == "__main__":
    main()
