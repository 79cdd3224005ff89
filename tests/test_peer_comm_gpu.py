"""The gradient exchange over peer memory (csrc/peer_comm.cu; replaces MpiAdamOptimizer.compute_gradients / apply_gradients,
spinup/utils/mpi_tf.py:45-80).  One GPU is enough for the protocol: three "ranks" are three comms on the same device whose slabs
are connected by pointer (ml4ca_peer_comm_connect_ptrs), one stream per rank -- the three kernels of a step run side by side and
wait for each other exactly as ranks on different GPUs do (tools/ppo_2gpu_check.py is the multi-GPU version, with CUDA IPC)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(world, n, dev):
    from ml4ca_b200 import _lib
    L = _lib.lib()
    comms, slabs = [], (ctypes.c_void_p * world)()
    for r in range(world):
        h = ctypes.c_void_p()
        _lib.check(L.ml4ca_peer_comm_create(r, world, n, dev.index or 0, ctypes.byref(h)), "ml4ca_peer_comm_create")
        p = ctypes.c_void_p()
        _lib.check(L.ml4ca_peer_comm_slab(h, ctypes.byref(p)))
        comms.append(h)
        slabs[r] = p.value
    for h in comms:
        _lib.check(L.ml4ca_peer_comm_connect_ptrs(h, slabs), "ml4ca_peer_comm_connect_ptrs")
    return comms


def _status(h):
    from ml4ca_b200 import _lib
    a, b = ctypes.c_int32(), ctypes.c_int32()
    _lib.check(_lib.lib().ml4ca_peer_comm_status(h, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def test_sum_is_identical_on_every_rank_and_in_rank_order(cuda_device):
    from ml4ca_b200 import _lib
    L, dev, world, n = _lib.lib(), cuda_device, 3, 10135
    comms = _make(world, n, dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    g = torch.Generator(device=dev); g.manual_seed(5)
    for step in range(7):                       # both halves of the slab, several times over
        bufs = [torch.randn(n, device=dev, generator=g) * 10.0 ** (step - 3) for _ in range(world)]
        stats = [torch.randn(8, device=dev, generator=g, dtype=torch.float64) for _ in range(world)]
        want = torch.zeros(n, device=dev)
        for r in range(world):                  # rank order, fp32
            src = bufs[r].clone()
            src[n - 8:n - 3] = stats[r][:5].float()
            want = want + src
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                _lib.check(L.ml4ca_peer_allreduce(comms[r], _lib.ptr(bufs[r]), n, _lib.ptr(stats[r]), n - 8, 5, None, 0,
                                                  _lib.current_stream()), "ml4ca_peer_allreduce")
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(bufs[r], want), (step, r)
    for h in comms:
        assert _status(h) == (7, 0)
        _lib.check(L.ml4ca_peer_comm_destroy(h))


def test_fused_exchange_and_adam_equals_the_separate_kernels(cuda_device):
    """ml4ca_adam_step_peer == rank-ordered sum, then ml4ca_adam_step_dev: parameters, moments, the statistics of iteration 0
    and the KL stop (ppo.py:268-271) bit for bit; a pass behind the stop is skipped on every rank alike."""
    from ml4ca_b200 import _lib
    L, dev, world, P = _lib.lib(), cuda_device, 2, 5000
    n, lo, hi = P + 8, 0, 3000
    comms = _make(world, n, dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    g = torch.Generator(device=dev); g.manual_seed(9)
    p0 = torch.randn(P, device=dev, generator=g)
    st = lambda: dict(p=p0.clone(), m1=torch.zeros(P, device=dev), m2=torch.zeros(P, device=dev),
                      ctl=torch.zeros(12, dtype=torch.int32, device=dev))
    ranks, ref = [st() for _ in range(world)], st()
    for s in ranks + [ref]:
        _lib.check(L.ml4ca_ppo_ctl_begin(_lib.ptr(s["ctl"]), None))
    count, limit = 1000.0, 0.015
    for it in range(4):
        flats = [torch.randn(n, device=dev, generator=g) for _ in range(world)]
        stats = [torch.rand(8, device=dev, generator=g, dtype=torch.float64) for _ in range(world)]
        for s_ in stats:
            s_[2] = 2.0 if it < 2 else 10.0      # rank-summed approx-KL / count: 0.004, 0.004, 0.02 > limit
        total = torch.zeros(n, device=dev)
        for r in range(world):
            src = flats[r].clone()
            src[P:P + 5] = stats[r][:5].float()
            total = total + src
        # reference: the separate Adam kernel on the summed buffer
        _lib.check(L.ml4ca_adam_step_dev(hi - lo, _lib.ptr(ref["p"][lo:hi]), _lib.ptr(total[lo:hi]), _lib.ptr(ref["m1"][lo:hi]),
                                         _lib.ptr(ref["m2"][lo:hi]), 3e-4, 0.9, 0.999, 1e-8, 1.0 / count, 0, it,
                                         _lib.ptr(total[P:P + 5]), count, limit, _lib.ptr(ref["ctl"]), None), "ml4ca_adam_step_dev")
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                s = ranks[r]
                _lib.check(L.ml4ca_adam_step_peer(comms[r], _lib.ptr(flats[r]), n, _lib.ptr(stats[r]), P, lo, hi, _lib.ptr(s["p"]),
                                                  _lib.ptr(s["m1"]), _lib.ptr(s["m2"]), 3e-4, 0.9, 0.999, 1e-8, 1.0 / count, 0, it,
                                                  count, limit, _lib.ptr(s["ctl"]), _lib.current_stream()), "ml4ca_adam_step_peer")
        torch.cuda.synchronize()
        for r in range(world):
            for k in ("p", "m1", "m2", "ctl"):
                assert torch.equal(ranks[r][k], ref[k]), (it, r, k)
    ctl = ref["ctl"].cpu().numpy()
    assert ctl[0] == 1 and ctl[1] == 2          # the KL sum of iteration 2 exceeded the limit: iteration 3 was skipped everywhere
    assert torch.equal(ref["p"][hi:], p0[hi:])
    for h in comms:
        assert _status(h) == (3, 0)             # three exchanges ran, the skipped pass did not count
        _lib.check(L.ml4ca_peer_comm_destroy(h))


def test_arguments_are_checked(cuda_device):
    from ml4ca_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.ml4ca_peer_comm_create(0, 1, 16, 0, ctypes.byref(h)) != 0          # a single rank has nothing to exchange
    assert L.ml4ca_peer_comm_create(2, 2, 16, 0, ctypes.byref(h)) != 0
    _lib.check(L.ml4ca_peer_comm_create(0, 2, 16, 0, ctypes.byref(h)))
    buf = torch.zeros(32, device=cuda_device)
    assert L.ml4ca_peer_allreduce(h, _lib.ptr(buf), 16, None, 0, 0, None, 0, None) != 0   # not connected
    _lib.check(L.ml4ca_peer_comm_destroy(h))


def test_a_silent_peer_times_out_instead_of_hanging(cuda_device):
    """Only rank 0 of two launches its exchange: the kernel gives up after ~5 s, counts the timeout and returns; later steps do
    not wait again."""
    import time
    from ml4ca_b200 import _lib
    L = _lib.lib()
    comms = _make(2, 64, cuda_device)
    buf = torch.ones(64, device=cuda_device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(L.ml4ca_peer_allreduce(comms[0], _lib.ptr(buf), 64, None, 0, 0, None, 0, None))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    assert 1.0 < t1 - t0 < 20.0
    assert _status(comms[0]) == (1, 1)
    _lib.check(L.ml4ca_peer_allreduce(comms[0], _lib.ptr(buf), 64, None, 0, 0, None, 0, None))
    torch.cuda.synchronize()
    assert time.perf_counter() - t1 < 1.0 and _status(comms[0]) == (2, 1)
    for h in comms:
        _lib.check(L.ml4ca_peer_comm_destroy(h))
