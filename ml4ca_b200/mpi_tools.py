"""Collectives of the PPO path: the reference's MPI helpers on torch.distributed (NCCL over NVLink on GPUs).

Mirrors /root/reference/src/rl/windows_workspace/spinup/utils/mpi_tools.py (proc_id, num_procs, mpi_avg,
mpi_statistics_scalar :43-93) and mpi_tf.py (sync_all_params :24, the gradient average inside
MpiAdamOptimizer.compute_gradients :45-70).  One process per GPU; the environments shard by global
index with no data-path collective -- these calls only carry gradients (one flat fp32 buffer per
optimizer step) and a handful of scalars.  With a single process every function is the identity.
The reference's per-step parameter Bcast (mpi_tf.py:72-80) is dropped: every rank applies the same
averaged gradient deterministically, so one broadcast at start-up (sync_all_params) suffices.
"""
import torch
import torch.distributed as dist


def _on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _comm_device(t=None):
    """Where a scalar must live for the collective: NCCL reduces CUDA tensors only (gloo takes CPU tensors)."""
    if t is not None and t.is_cuda:
        return t.device
    if _on() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def proc_id():
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def num_procs():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def shard_bounds(n_global, rank=None, world=None):
    """Global env index range [lo, hi) of a rank: contiguous blocks, remainder spread over the first ranks."""
    rank = proc_id() if rank is None else rank
    world = num_procs() if world is None else world
    base, rem = divmod(int(n_global), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t):
    """In-place sum over ranks (mpi_tools.allreduce :47-52)."""
    if _on():
        if t.is_contiguous():
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        else:                       # a strided view (e.g. columns of a statistics table): reduce a packed copy, write it back
            c = t.contiguous()
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            t.copy_(c)
    return t


def mpi_avg(x):
    """Average a scalar / tensor over ranks (:67-69)."""
    if not _on():
        return x
    t = x if torch.is_tensor(x) else torch.tensor(float(x), dtype=torch.float64)
    src = t.device
    t = t.to(_comm_device(t), copy=True)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    t /= num_procs()
    return t.to(src) if torch.is_tensor(x) else float(t.item())


def average_gradients_(flat_grad):
    """MpiAdamOptimizer.compute_gradients (mpi_tf.py:59-62): Allreduce(SUM) of the flat gradient, divided by the
    number of ranks -- ONE collective per optimizer step on one contiguous buffer."""
    if _on():
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad /= num_procs()
    return flat_grad


def sync_all_params(flat_params, root=0):
    """sync_all_params (mpi_tf.py:24-27): broadcast rank 0's parameters once."""
    if _on():
        dist.broadcast(flat_params, src=root)
    return flat_params


class PeerComm(object):
    """The gradient exchange of MpiAdamOptimizer (mpi_tf.py:45-80) over NVLink peer memory: ml4ca_peer_allreduce /
    ml4ca_adam_step_peer (csrc/peer_comm.cu) sum a small flat buffer over the GPUs of one node inside ONE kernel that also
    applies the Adam step -- no NCCL call, no separate scale / copy kernels.  ``PeerComm.create`` returns None (callers keep
    the NCCL all-reduce) when there is a single rank, the backend is not NCCL, ML4CA_PEER_COMM=0, or any rank fails to map its
    peers' memory (ranks on different nodes, no peer access): the decision is taken collectively."""

    def __init__(self, handle, device):
        self._handle, self.device = handle, device

    @classmethod
    def create(cls, max_floats, device):
        import ctypes, os
        from . import _lib
        if not _on() or dist.get_backend() != "nccl" or os.environ.get("ML4CA_PEER_COMM", "1") == "0":
            return None
        if num_procs() > 16:
            return None
        device = torch.device(device)
        L, h, ok = _lib.lib(), ctypes.c_void_p(), 1
        mine = torch.zeros(64, dtype=torch.uint8)
        with torch.cuda.device(device):
            if L.ml4ca_peer_comm_create(proc_id(), num_procs(), int(max_floats), device.index, ctypes.byref(h)) != 0:
                ok, h = 0, None
            elif L.ml4ca_peer_comm_export(h, mine.data_ptr()) != 0:
                ok = 0
            every = [torch.zeros(64, dtype=torch.uint8, device=device) for _ in range(num_procs())]
            dist.all_gather(every, mine.to(device))
            if ok and L.ml4ca_peer_comm_connect(h, torch.cat(every).cpu().contiguous().data_ptr()) != 0:
                ok = 0
            agree = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)       # also the barrier: every slab is mapped before anyone uses it
            if int(agree.item()) == 0:
                if h is not None:
                    L.ml4ca_peer_comm_destroy(h)
                if proc_id() == 0:
                    import sys
                    print("ml4ca_b200: peer-memory gradient exchange unavailable (%s); using NCCL"
                          % _lib.lib().ml4ca_last_error().decode("utf-8", "replace"), file=sys.stderr)
                return None
        return cls(h, device)

    def status(self):
        """(steps completed, waits given up).  A non-zero second number means a peer did not answer within 5 s."""
        import ctypes
        from . import _lib
        a, b = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(_lib.lib().ml4ca_peer_comm_status(self._handle, ctypes.byref(a), ctypes.byref(b)), "ml4ca_peer_comm_status")
        return a.value, b.value

    def close(self):
        if self._handle is not None:
            from . import _lib
            if dist.is_initialized():
                torch.cuda.synchronize(self.device)
                dist.barrier()                                  # nobody unmaps a slab a peer's kernel may still read
            _lib.lib().ml4ca_peer_comm_destroy(self._handle)
            self._handle = None


def statistics_from_sums(sums3):
    """[sum, sum of squares, count] (already reduced over ranks) -> (mean, std), population std like
    mpi_statistics_scalar (:85-89: sqrt(sum((x - mean)^2) / n))."""
    s, q, n = float(sums3[0]), float(sums3[1]), float(sums3[2])
    mean = s / n
    var = max(q / n - mean * mean, 0.0)
    return mean, var ** 0.5


def mpi_statistics_scalar(x, with_min_and_max=False):
    """mpi_statistics_scalar (:71-93) for a tensor on any device."""
    x = torch.as_tensor(x).to(torch.float64).reshape(-1)
    x = x.to(_comm_device(x))
    sums = torch.stack([x.sum(), (x * x).sum(), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
    allreduce_sum_(sums)
    mean, std = statistics_from_sums(sums.tolist())
    if with_min_and_max:
        lo, hi = x.min().clone(), x.max().clone()
        if _on():
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        return mean, std, float(lo), float(hi)
    return mean, std
