"""Experiment: BASELINE configs[3] rollout as two half-batches on two streams (policy kernel of one half next to the env kernel of
the other) against one batch on one stream.  Tuning tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ml4ca_b200 as M
from ml4ca_b200 import _lib

n = int(os.environ.get("N", 1 << 23))
P = int(os.environ.get("PARTS", 2))
T = 8
dev = torch.device("cuda", 0)
L = _lib.lib()
ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", device=dev, seed=3)


def make(parts):
    envs, bufs = [], []
    for k in range(parts):
        e = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=n // parts, device=dev, seed=4,
                          auto_reset=True, env_id_offset=k * (n // parts))
        e.reset(fraction=0.8)
        b = M.TrajectoryBuffer(9, 7, T, n // parts, device=dev)
        b._obs_rows[0].copy_(e._obs)
        envs.append(e), bufs.append(b)
    return envs, bufs


def run(envs, bufs, streams, start):
    for t in range(T):
        for e, b, s in zip(envs, bufs, streams):
            with torch.cuda.stream(s):
                st = _lib.current_stream()
                rows = b._obs_rows
                _lib.check(L.ml4ca_policy_forward(ac._handle, e.num_envs, _lib.ptr(rows[t]), 1, start + t, 0, e._cfg.env_id_offset,
                                                  _lib.ptr(b.act_buf[t]), _lib.ptr(b.val_buf[t]), _lib.ptr(b.logp_buf[t]), None, st))
                e.step_into(b.act_buf[t], rows[t + 1], b.rew_buf[t], b.done_buf[t])
    for e, b, s in zip(envs, bufs, streams):
        with torch.cuda.stream(s):
            b._obs_rows[0].copy_(b._obs_rows[T])


for parts in (1, P):
    envs, bufs = make(parts)
    streams = [torch.cuda.Stream(dev) for _ in range(parts)]
    torch.cuda.synchronize()
    for i in range(2):
        run(envs, bufs, streams, i * T)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for i in range(4):
        run(envs, bufs, streams, (2 + i) * T)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (4 * T)
    print("n %d in %d part(s) on %d stream(s): %.4f ms per step = %.2f G env-steps/s" % (n, parts, parts, ms, n / ms / 1e6))
