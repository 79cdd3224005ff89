"""CPU tests of the PPO plumbing: GAE oracle vs the reference formulation's golden vector, and the
collective helpers on a 2-process gloo group (the N > 1 host path; NCCL is the same code on GPUs)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden
from oracle import ppo_oracle as PO


def test_gae_oracle_matches_reference_formulation():
    g = golden("gae.npz")
    adv, ret = PO.finish_path(g['rews'], g['vals'], float(g['last_val']), float(g['gamma']), float(g['lam']))
    np.testing.assert_allclose(adv, g['adv'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(ret, g['ret'], rtol=0, atol=2e-6)
    # the batched form with no flags is the same single trajectory
    T = len(g['rews'])
    val = np.append(g['vals'], g['last_val'])[:, None]
    a2, r2 = PO.gae_batched(g['rews'][:, None], val, np.zeros((T, 1), dtype=np.uint8), float(g['gamma']), float(g['lam']))
    np.testing.assert_allclose(a2[:, 0], g['adv'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(r2[:, 0], g['ret'], rtol=0, atol=2e-6)


def test_shard_bounds_cover_every_env_once():
    from ml4ca_b200 import mpi_tools
    for n, w in [(16, 1), (17, 2), (1 << 24, 8), (5, 8)]:
        spans = [mpi_tools.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml4ca_b200 import mpi_tools
    rng = np.random.default_rng(100 + rank)
    x = torch.as_tensor(rng.normal(loc=rank, size=1000 + 10 * rank))
    mean, std, lo, hi = mpi_tools.mpi_statistics_scalar(x, with_min_and_max=True)
    grad = torch.full((7,), float(rank + 1))
    mpi_tools.average_gradients_(grad)
    params = torch.full((5,), float(rank))
    mpi_tools.sync_all_params(params)
    kl = mpi_tools.mpi_avg(0.01 * (rank + 1))
    table = torch.arange(15, dtype=torch.float64).reshape(3, 5) * (rank + 1)     # the epoch-statistics table of run_epochs:
    mpi_tools.allreduce_sum_(table[:, 0:3])                                      # only its sum columns are rank-summed (strided view)
    out[rank] = dict(table=table.tolist(), mean=mean, std=std, lo=lo, hi=hi, grad=grad.tolist(), params=params.tolist(), kl=kl,
                     ids=(mpi_tools.proc_id(), mpi_tools.num_procs()), x=x.numpy())
    dist.destroy_process_group()


def test_collectives_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    allx = np.concatenate([out[0]['x'], out[1]['x']])
    for r in range(2):
        o = out[r]
        assert o['ids'] == (r, 2)
        assert abs(o['mean'] - allx.mean()) < 1e-12 and abs(o['std'] - allx.std()) < 1e-12
        assert o['lo'] == allx.min() and o['hi'] == allx.max()
        assert o['grad'] == [1.5] * 7                 # mean of 1 and 2: Allreduce(SUM) / num_procs (mpi_tf.py:59-62)
        assert o['params'] == [0.0] * 5               # broadcast from rank 0 (mpi_tf.py:24-27)
        assert abs(o['kl'] - 0.015) < 1e-15           # mpi_avg (mpi_tools.py:67-69)
        want = np.arange(15, dtype=np.float64).reshape(3, 5) * (r + 1)
        want[:, 0:3] = np.arange(15, dtype=np.float64).reshape(3, 5)[:, 0:3] * 3
        assert o['table'] == want.tolist()


def test_gradient_oracle_matches_torch_autograd():
    """oracle.ppo_oracle.ppo_gradients (analytic, float64) == torch autograd of the losses written as in ppo.py:234-237
    with core.py's gaussian_likelihood: pins the checker the GPU update kernel is compared with."""
    from oracle import mlp_oracle as MO
    dims = dict(obs_dim=9, act_dim=7, hidden=64, n_hidden=2)
    for actn in ("leaky_relu", "tanh"):
        rng = np.random.default_rng(0)
        N = 300
        base = MO.glorot_params(dims, 3).astype(np.float64)
        flat = base + rng.normal(size=base.size) * 0.05
        obs, act = rng.normal(size=(N, 9)), rng.normal(size=(N, 7))
        adv, ret = rng.normal(size=N), rng.normal(size=N) * 3
        fo = MO.forward(flat * 1.02, dims, obs.T, actn)
        lpo = MO.gaussian_likelihood(act, fo["mu"].T, fo["log_std"])
        g, info = PO.ppo_gradients(flat, dims, obs, act, adv, ret, lpo, 0.2, actn)
        p = torch.tensor(flat, requires_grad=True)
        pos = [0]

        def take(*shape):
            k = int(np.prod(shape))
            t = p[pos[0]:pos[0] + k].reshape(*shape)
            pos[0] += k
            return t

        def net(o):
            sizes = [9, 64, 64, o]
            return [(take(sizes[i], sizes[i + 1]), take(sizes[i + 1])) for i in range(3)]

        pi, ls, vl = net(7), take(7), net(1)
        f = (lambda z: torch.where(z > 0, z, 0.2 * z)) if actn == "leaky_relu" else torch.tanh

        def mlp(x, L):
            for W, b in L[:-1]:
                x = f(x @ W + b)
            return x @ L[-1][0] + L[-1][1]

        x = torch.tensor(obs)
        mu, v = mlp(x, pi), mlp(x, vl)[:, 0]
        logp = (-0.5 * (((torch.tensor(act) - mu) / (torch.exp(ls) + 1e-8)) ** 2 + 2 * ls + np.log(2 * np.pi))).sum(1)
        ratio, a = torch.exp(logp - torch.tensor(lpo)), torch.tensor(adv)
        pi_loss = -torch.minimum(ratio * a, torch.where(a > 0, 1.2 * a, 0.8 * a)).mean()
        v_loss = ((torch.tensor(ret) - v) ** 2).mean()
        (pi_loss + v_loss).backward()
        np.testing.assert_allclose(p.grad.numpy(), g, rtol=0, atol=1e-12)
        assert abs(pi_loss.item() - info["pi_loss"]) < 1e-12 and abs(v_loss.item() - info["v_loss"]) < 1e-12
