"""Timing of the PPO update kernels and one small end-to-end training run (tuning tool)."""
import sys, time
sys.path.insert(0, '.')
import torch
import ml4ca_b200 as M

dev = torch.device('cuda', 0)
T, n = 64, 1 << 16
ac = M.ActorCritic(9, 7, (64, 64), 'leaky_relu', device=dev, seed=1)
g = torch.Generator(device=dev); g.manual_seed(0)
data = (torch.randn(T, 9, n, device=dev, generator=g), torch.randn(T, 7, n, device=dev, generator=g),
        torch.randn(T, n, device=dev, generator=g), torch.randn(T, n, device=dev, generator=g),
        torch.randn(T, n, device=dev, generator=g) - 9.0)
upd = M.PPOUpdater(ac)
for net in (0, 1):
    upd._grad(net, data, T, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        upd._grad(net, data, T, n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('net %d: %.3f ms per pass over %d samples -> %.1f M sample-passes/s, %.1f TFLOP/s fp32 (2 x 15.6 k FMA)' % (
        net, ms, T * n, T * n / ms / 1e3, T * n * 15600 * 2 / ms / 1e9))
env = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=1 << 14, device=dev, seed=3, auto_reset=True)
t0 = time.perf_counter()
ac2, hist = M.ppo(env, steps_per_epoch=100, epochs=3, train_pi_iters=20, train_v_iters=20, seed=3)
torch.cuda.synchronize()
print('3 epochs of 16 Ki envs x 100 steps: %.2f s' % (time.perf_counter() - t0))
for h in hist:
    print({k: (round(v, 5) if isinstance(v, float) else v) for k, v in h.items()})
