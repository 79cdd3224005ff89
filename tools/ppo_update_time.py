"""Device time of one graph-replayed PPO update (80 + 80 iterations) under torchrun: the gradient exchange over NVLink peer memory
(default) against NCCL (ML4CA_PEER_COMM=0), at the bench batch (16 Ki envs x 400 steps per rank) and the reference's (4 x 400).
HIDDEN=80,80,80 times the reference's own network instead of the 64 x 64 of the bench config (ML4CA_PPO_FP32=1: the fp32 kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import ml4ca_b200 as M

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
for n in [int(x) for x in os.environ.get("ENVS", "16384,4").split(",")]:
    T = 400
    hidden = tuple(int(x) for x in os.environ.get("HIDDEN", "64,64").split(","))
    ac = M.ActorCritic(9, 7, hidden, "leaky_relu", device=dev, seed=4)
    buf = M.TrajectoryBuffer(9, 7, T, n, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(7 + rank)
    buf.obs_buf.normal_(generator=g); buf.adv_buf.normal_(generator=g); buf.ret_buf.normal_(generator=g)
    o = buf.obs_buf.permute(1, 0, 2).reshape(9, -1).contiguous()
    a, _, lp = ac.step(o, deterministic=False, step=0)
    buf.act_buf.copy_(a.reshape(7, T, n).permute(1, 0, 2)); buf.logp_buf.copy_(lp.reshape(T, n))
    upd = M.PPOUpdater(ac, target_kl=1e9)          # never stops: all 80 + 80 iterations run
    buf.get = lambda: (buf.obs_buf, buf.act_buf, buf.adv_buf, buf.ret_buf, buf.logp_buf)
    for _ in range(2):
        upd.update(buf, graph=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        upd.update(buf, graph=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    if rank == 0:
        print("%d rank(s), hidden %s, exchange %s, %5d envs x 400: update %.3f ms = %.1f us per iteration = %.1f M sample-passes/s%s" % (
            world, hidden, "none" if world == 1 else ("peer" if upd.peer is not None else "NCCL"), n, ms, ms / 160 * 1e3,
            162 * n * T / ms / 1e3,
            "" if upd.peer is None else "  peer status %s" % (upd.peer.status(),)))
if world > 1:
    dist.destroy_process_group()
