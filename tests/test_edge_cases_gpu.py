"""Empty and degenerate inputs through the C ABI: n = 0 is a no-op that launches nothing, NULL buffers are refused with
ML4CA_ERR_INVALID and a message (the reference raises AssertionError on bad arguments, customEnv.py:35,228,238), a
single environment reproduces the reference's call shape, and a full-size QP batch satisfies the first-order
conditions everywhere (size-independent property at BASELINE config 1 scale x 256)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import qp_oracle as QO

pytestmark = pytest.mark.gpu


def test_empty_batches_are_no_ops(cuda_device):
    from ml4ca_b200 import _lib
    L = _lib.lib()
    buf = torch.zeros(64, device=cuda_device)                # non-NULL dummy for every pointer argument
    p, st = _lib.ptr(buf), _lib.current_stream()
    before = _lib.launch_count()
    assert L.ml4ca_qp_solve(0, p, p, p, p, st) == 0
    assert L.ml4ca_pinv_pid(0, p, p, p, p, p, p, p, st) == 0
    assert L.ml4ca_pinv_allocate(0, p, p, p, st) == 0
    assert L.ml4ca_error_frame(0, p, p, p, st) == 0
    assert L.ml4ca_gae(0, 5, p, p, p, None, 1, 0.99, 0.97, p, p, st) == 0
    assert L.ml4ca_gae(7, 0, p, p, p, None, 1, 0.99, 0.97, p, p, st) == 0
    import ml4ca_b200 as M
    ac = M.ActorCritic(9, 7, (64, 64), "leaky_relu", device=cuda_device)
    before = _lib.launch_count()
    assert L.ml4ca_policy_forward(ac._handle, 0, p, 0, 0, 0, 0, p, p, p, None, st) == 0
    grad, stats = torch.ones(ac.num_params + 8, device=cuda_device), torch.ones(8, dtype=torch.float64, device=cuda_device)
    assert L.ml4ca_ppo_grad(ac._handle, 0, 0, 3, p, p, p, p, p, 0.2, _lib.ptr(grad), _lib.ptr(stats), st) == 0
    assert L.ml4ca_trpo_kl_grad(ac._handle, 0, 3, p, p, p, _lib.ptr(grad), _lib.ptr(stats), st) == 0
    assert L.ml4ca_trpo_policy_mu(ac._handle, 0, 3, p, p, st) == 0
    torch.cuda.synchronize()
    assert _lib.launch_count() == before                     # nothing launched
    assert float(grad[:ac.num_params].abs().max()) == 0 and float(stats.abs().max()) == 0   # an empty pass still zeroes its outputs


def test_null_buffers_are_refused(cuda_device):
    from ml4ca_b200 import _lib
    L = _lib.lib()
    buf = torch.zeros(64, device=cuda_device)
    p, st = _lib.ptr(buf), _lib.current_stream()
    assert L.ml4ca_qp_solve(4, None, p, p, p, st) == -1
    assert b"ml4ca_qp_solve" in L.ml4ca_last_error()
    assert L.ml4ca_pinv_pid(4, p, p, p, p, p, None, p, st) == -1
    assert L.ml4ca_qp_solve(-1, p, p, p, p, st) == -1
    assert L.ml4ca_env_step(None, p, p, p, p, st) == -1
    with pytest.raises(_lib.Ml4caError):
        _lib.check(L.ml4ca_env_reset(None, None, 0.8, p, st), "ml4ca_env_reset")
    handle = ctypes.c_void_p()
    cfg = _lib.EnvCfg()
    assert L.ml4ca_env_cfg_default(3, 1, 1, ctypes.byref(cfg)) == 0
    assert L.ml4ca_env_create(ctypes.byref(cfg), -5, 0, ctypes.byref(handle)) == -1
    cfg.max_ep_len = 0
    assert L.ml4ca_env_create(ctypes.byref(cfg), 16, 0, ctypes.byref(handle)) == -1


def test_qp_first_order_conditions_at_full_size(cuda_device):
    """1 Mi allocations of the config-1 demand law: every reported success is feasible for the reference problem
    (bounds, rate limits, slack bounds) and the |x| < 0.01 clean-up is applied everywhere -- vectorised over the batch."""
    import ml4ca_b200 as M
    n = 1 << 20
    tau, prev = QO.synth_batch(4096, seed=0)
    reps = n // 4096
    tau, prev = np.tile(tau, reps).astype(np.float32), np.tile(prev, reps).astype(np.float32)
    ta = M.QPTA(num_envs=n, device=cuda_device)
    ta.previous_thruster_state = prev.astype(np.float64)
    x, ok = ta.solve_QP(torch.as_tensor(tau, device=cuda_device))
    x, ok = x.cpu().numpy().astype(np.float64), ok.cpu().numpy()
    # periodic input -> periodic output: every tile of 4096 demands got the same answer (no cross-talk between lanes / CTAs)
    x3 = x.reshape(8, reps, 4096)
    assert np.array_equal(x3, np.broadcast_to(x3[:, :1], x3.shape))
    assert np.array_equal(ok.reshape(reps, 4096), np.broadcast_to(ok[:4096], (reps, 4096)))
    x, ok, tau64, prev64 = x[:, :4096], ok[:4096], tau[:, :4096].astype(np.float64), prev[:, :4096].astype(np.float64)
    assert ok[: int(0.9 * 4096)].mean() > 0.95
    s = ok
    C = QO.C
    fmax = np.array(C.F_MAX)[:, None]
    df = np.array(C.QP_DF)[:, None]
    da = np.array(C.QP_DA)[:, None]
    tol = 2e-4
    assert (np.abs(x[0:3, s]) <= fmax + tol).all()                                     # qp_allocator.py:196-200
    assert (np.abs(x[3:5, s]) <= C.QP_ALPHA_BOUND + tol).all()
    assert (np.abs(x[5:8, s]) <= C.QP_SLACK_BOUND + tol).all()
    assert (np.abs(x[0:3, s] - prev64[0:3, s]) <= df + 0.01 + tol).all()               # :164-169 (clean-up may move f by < 0.01)
    assert (np.abs(x[3:5, s] - prev64[3:5, s]) <= da + 0.01 + tol).all()               # :172-175
    raw = s & (x[0:5] != 0).all(axis=0)                                                # no force / angle touched by the clean-up
    res = QO.wrench_rows(x[0:3, raw], x[3:5, raw]) - tau64[:, raw]                     # B(alpha) f - s = tau, :156-158
    assert raw.mean() > 0.5 and np.abs(res - x[5:8, raw]).max() < 0.01 + tol           # (a slack itself may be cleaned to 0)
    assert not ((np.abs(x) < C.QP_CLEAN_EPS) & (x != 0)).any()                         # :232
