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
    out[rank] = dict(mean=mean, std=std, lo=lo, hi=hi, grad=grad.tolist(), params=params.tolist(), kl=kl,
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
