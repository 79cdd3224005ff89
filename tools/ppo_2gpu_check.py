"""N-rank check of the PPO path (run under torchrun): every rank must end with identical parameters, and the rank-summed
gradient of the sharded batch must equal the single-process gradient of the whole batch.  The gradient exchange runs over NVLink
peer memory (csrc/peer_comm.cu) unless ML4CA_PEER_COMM=0 selects the NCCL all-reduce; the parameter digest printed at the end
lets the two be compared (bit-identical for 2 ranks).  EPOCH_ENVS=16384 adds the epoch time at the bench size."""
import os, sys
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import ml4ca_b200 as M
from ml4ca_b200 import mpi_tools

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n_glob, T = 8192, 8
lo, hi = mpi_tools.shard_bounds(n_glob)
g = torch.Generator(device=dev); g.manual_seed(1)                     # same global batch on every rank
full = (torch.randn(T, 9, n_glob, device=dev, generator=g), torch.randn(T, 7, n_glob, device=dev, generator=g),
        torch.randn(T, n_glob, device=dev, generator=g), torch.randn(T, n_glob, device=dev, generator=g),
        torch.randn(T, n_glob, device=dev, generator=g) - 9.0)
shard = tuple(x[..., lo:hi].contiguous() for x in full)
ac = M.ActorCritic(9, 7, (64, 64), 'leaky_relu', device=dev, seed=4)     # same seed -> same init; sync anyway
mpi_tools.sync_all_params(ac.parameters()); ac.refresh()
upd = M.PPOUpdater(ac)
print("rank %d: gradient exchange over %s" % (rank, "NVLink peer memory" if upd.peer is not None else "NCCL"))
if upd.peer is not None:                     # the exchange kernel against NCCL on a random buffer
    P8 = ac.num_params + 8
    g2 = torch.Generator(device=dev); g2.manual_seed(100 + rank)
    for rep in range(5):
        a = torch.randn(P8, device=dev, generator=g2)
        b = a.clone()
        L_ = __import__("ml4ca_b200")._lib
        L_.check(L_.lib().ml4ca_peer_allreduce(upd.peer._handle, L_.ptr(a), P8, None, 0, 0, None, 0, L_.current_stream()))
        dist.all_reduce(b)
        err = (a - b).abs().max().item()
        assert err <= (0.0 if world == 2 else 1e-5), err
    def per_call(fn, reps=300):
        for _ in range(20):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3
    t_peer = per_call(lambda: L_.lib().ml4ca_peer_allreduce(upd.peer._handle, L_.ptr(a), P8, None, 0, 0, None, 0, L_.current_stream()))
    t_nccl = per_call(lambda: dist.all_reduce(b))
    print("rank %d: %d floats, back-to-back calls: peer kernel %.1f us, NCCL all_reduce %.1f us" % (rank, P8, t_peer, t_nccl))
    print("rank %d: ml4ca_peer_allreduce == NCCL all_reduce on 5 random buffers (max diff %.1e), status %s" % (rank, err, upd.peer.status()))
s, c = upd._grad(0, shard, T, hi - lo)
g_dist = upd.flat[:ac.num_params].clone()
dist.destroy_process_group() if False else None
# single-process reference on the whole batch (collectives are a no-op only when uninitialised: compute by hand)
import ml4ca_b200._lib as L
flat = torch.zeros_like(upd.flat); stats = torch.zeros(8, dtype=torch.float64, device=dev)
L.check(L.lib().ml4ca_ppo_grad(ac._handle, 0, n_glob, T, *[L.ptr(x) for x in full[:3]], L.ptr(full[3]), L.ptr(full[4]), 0.2,
                               L.ptr(flat), L.ptr(stats), L.current_stream()))
err = (g_dist - flat[:ac.num_params]).abs().max().item() / flat[:ac.num_params].abs().max().item()
print("rank %d: count %.0f, rank-summed gradient vs whole-batch gradient: rel err %.2e" % (rank, c, err))
assert c == n_glob * T and err < 1e-5
env = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=hi - lo, device=dev, seed=3,
                    auto_reset=True, env_id_offset=lo)
ac2, hist = M.ppo(env, steps_per_epoch=50, epochs=2, train_pi_iters=5, train_v_iters=5, seed=3)
p = ac2.parameters().clone()
ref = p.clone(); dist.broadcast(ref, src=0)
assert torch.equal(p, ref), "parameters diverged between ranks"
print("rank %d: 2 PPO epochs over the collective ok, parameters identical on all ranks; last %s" % (rank, {k: hist[-1][k] for k in ("KL", "LossV", "StopIter")}))
# the same with rollout AND update replayed from CUDA graphs (the NCCL all-reduce of the flat gradient is captured with them)
import time
env = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=hi - lo, device=dev, seed=3,
                    auto_reset=True, env_id_offset=lo)
marks = []
def mark(info):
    torch.cuda.synchronize(); marks.append(time.perf_counter())
ac3, hist3 = M.ppo(env, steps_per_epoch=50, epochs=5, train_pi_iters=5, train_v_iters=5, seed=3, graph=True, logger=mark)
p = ac3.parameters().clone()
ref = p.clone(); dist.broadcast(ref, src=0)
assert torch.equal(p, ref), "parameters diverged between ranks (graph update)"
print("rank %d: 5 PPO epochs with graph-replayed update over the collective ok, parameters identical; epoch %.2f ms; last %s; digest %s" % (
    rank, 1e3 * (marks[-1] - marks[-2]), {k: hist3[-1][k] for k in ("KL", "LossV", "StopIter")},
    p.double().sum().item().hex()))
ne = int(os.environ.get("EPOCH_ENVS", 0))
if ne:
    env = M.RevoltFinal(M.StandInHull(), extended_state=True, cont_ang=True, num_envs=ne, device=dev, seed=5, auto_reset=True,
                        env_id_offset=rank * ne)
    marks = []
    _, hist4 = M.ppo(env, steps_per_epoch=400, epochs=6, seed=5, graph=True, logger=mark)
    dts = sorted(b - a for a, b in zip(marks[1:], marks[2:]))
    if rank == 0:
        print("%d ranks x %d envs x 400 steps: epoch %.2f ms (median of %d), StopIter %s" % (
            world, ne, 1e3 * dts[len(dts) // 2], len(dts), [h["StopIter"] for h in hist4]))
dist.destroy_process_group()
