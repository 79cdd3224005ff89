"""The per-demand solver of the CUDA allocator kernel (ml4ca_b200/csrc/qp_slsqp.cuh), compiled for the HOST by g++ as a
test harness (tests/host_harness/qp_host.cpp), against the reference's outputs on the full config-1 batch.

This is the same source every GPU thread runs (the header is shared between the __device__ and the host build), so the
path-following logic is pinned here without a GPU; tests/test_qp_gpu.py repeats the comparison through the C ABI on the
device.  The harness is test infrastructure: it is not part of libml4ca_b200.so and nothing in the product calls it."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, golden

sys.path.insert(0, os.path.join(ROOT, "tests"))
import qp_parity  # noqa: E402


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("qp_host") / "qp_host")
    src = os.path.join(ROOT, "tests", "host_harness", "qp_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, src], check=True)
    return exe


def run(exe, kind, tau, prev, tmp, weights=None, fuel=True):
    n = tau.shape[1]
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    with open(fin, "wb") as fh:
        np.array([n], dtype=np.int64).tofile(fh)
        np.ascontiguousarray(tau, dtype=np.float64).tofile(fh)
        np.ascontiguousarray(prev, dtype=np.float64).tofile(fh)
    extra = [] if weights is None else [repr(float(w)) for w in weights] + [str(int(fuel))]
    subprocess.run([exe, kind, fin, fout] + extra, check=True)
    raw = np.fromfile(fout, dtype=np.uint8)
    x = raw[:8 * n * 8].view(np.float64).reshape(8, n)
    rest = raw[8 * n * 8:].view(np.int32).reshape(3, n)
    return x, rest[0], rest[1], rest[2]


def test_float64_path_reproduces_the_reference_row_by_row(harness, tmp_path):
    g = golden("qp_config1.npz")
    assert g['tau'].shape[1] == 4096                                   # BASELINE configs[0] at its stated size
    x, mode, it, mask = run(harness, "double", g['tau'], g['prev'], str(tmp_path))
    cls, dist = qp_parity.classify(g, x, mode == 0, mask)
    counts = {k: int((cls == k).sum()) for k in ('A', 'B', 'C', 'FAIL', 'FLAG', 'X')}
    assert counts['FLAG'] == 0 and counts['X'] == 0, qp_parity.table(g, cls, dist, it)
    assert counts['A'] >= 0.99 * (counts['A'] + counts['B'] + counts['C']), counts
    assert counts['C'] <= 10, counts
    both = np.isin(cls, ['A', 'B', 'C'])
    np.testing.assert_array_equal(mask[both], g['mask_raw'][both])    # the reference's active set, exactly
    same_iter = (it[both] == g['slsqp_nit'][both]).mean()
    assert same_iter >= 0.98, same_iter                                # stops at the iteration the reference stops at


def test_lane_distributed_formulation_is_the_same_solver(harness, tmp_path):
    """csrc/qp_group.cuh (one demand over 8 lanes, the alternative kernel mapping) on its host backend: same flags, same
    iteration counts, same end points as the scalar formulation (sums are re-associated: 1e-9)."""
    g = golden("qp_config1.npz")
    xa, ma, ia, ka = run(harness, "double", g['tau'], g['prev'], str(tmp_path))
    xb, mb, ib, kb = run(harness, "group", g['tau'], g['prev'], str(tmp_path))
    np.testing.assert_array_equal(ma == 0, mb == 0)
    same = (ma == 0) & (ia == ib)
    assert same.sum() >= 0.995 * (ma == 0).sum()
    np.testing.assert_allclose(xb[:, same], xa[:, same], rtol=0, atol=1e-7)
    cls, dist = qp_parity.classify(g, xb, mb == 0, kb)
    assert int((cls == 'FLAG').sum()) == 0 and int((cls == 'X').sum()) == 0


def test_float32_cannot_follow_the_path(harness, tmp_path):
    """Why the kernel computes in float64: the reference's stopping tests compare |f - f0| with 1e-6, below the
    fp32 resolution of f ~ 10..200; an fp32 path overruns them."""
    g = golden("qp_config1.npz")
    n = 1024
    x, mode, it, mask = run(harness, "float", g['tau'][:, :n], g['prev'][:, :n], str(tmp_path))
    ref = g['success'][:n]
    assert ((mode == 0) != ref).mean() > 0.02                          # measured 0.058
    both = (mode == 0) & ref
    assert (it[both] == g['slsqp_nit'][:n][both]).mean() < 0.5         # measured 0.16


@pytest.mark.parametrize("case", ["nofuel", "noflick", "noang", "bare", "weighted"])
def test_objective_switches_host(harness, tmp_path, case):
    """reduce_fuel / reduce_flickering / reduce_angular and a diagonal weight_matrix (:108,116-150) against the reference
    called with those arguments (tests/golden/qp_switches.npz)."""
    g = golden("qp_switches.npz")
    x, mode, it, mask = run(harness, "double", g['tau'], g['prev'], str(tmp_path), g[case + '__weights'],
                            bool(g[case + '__fuel']))
    ok, ref = mode == 0, g[case + '__success']
    np.testing.assert_array_equal(ok, ref)
    d_ref = qp_parity._rel(x, g[case + '__x_raw'])
    d_exact = qp_parity._rel(x, g[case + '__x_raw_exact'])
    both = ok & ref
    literal = d_ref[both] <= 1e-5
    explained = literal | ((d_exact[both] <= 1e-5) & g[case + '__success_exact'][both])
    assert literal.mean() >= 0.97, (case, literal.mean())
    assert d_ref[both].max() <= 1e-3, (case, d_ref[both].max())                 # never another basin
    assert (~explained).sum() <= 2, (case, (~explained).sum())   # class C of qp_parity (no perturbation runs in this fixture)
