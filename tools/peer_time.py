"""Per-call device time of the gradient exchange inside a CUDA graph (as the PPO update replays it): ml4ca_peer_allreduce against
NCCL all_reduce on the same 57 KB buffer.  Run under torchrun, one rank per GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from ml4ca_b200 import _lib, mpi_tools

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = int(os.environ.get("FLOATS", 14343))
peer = mpi_tools.PeerComm.create(n, dev)
assert peer is not None
L = _lib.lib()
a = torch.randn(n, device=dev)
b = a.clone()
K = 200


def graph_of(fn):
    fn(); torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="relaxed"):
        for _ in range(K):
            fn()
    return g


def time_graph(g):
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * K) * 1e3


t_peer = time_graph(graph_of(lambda: _lib.check(L.ml4ca_peer_allreduce(peer._handle, _lib.ptr(a), n, None, 0, 0, None, 0, _lib.current_stream()))))
t_nccl = time_graph(graph_of(lambda: dist.all_reduce(b)))
print("rank %d of %d: %d floats inside a CUDA graph: peer kernel %.2f us per call, NCCL all_reduce %.2f us per call; status %s" % (
    rank, world, n, t_peer, t_nccl, peer.status()))
peer.close()
dist.destroy_process_group()
