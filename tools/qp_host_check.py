"""Run the host build of the CUDA allocator's per-demand solver (tests/host_harness/qp_host.cpp) over the config-1 golden
batch and print its agreement with the reference's outputs.   python tools/qp_host_check.py [float|double|mixed] [n]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(exe="/tmp/qp_host"):
    src = os.path.join(ROOT, "tests", "host_harness", "qp_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, src], check=True)
    return exe


def run(exe, kind, tau, prev):
    n = tau.shape[1]
    fin, fout = "/tmp/qp_host_in_%d.bin" % os.getpid(), "/tmp/qp_host_out_%d.bin" % os.getpid()
    with open(fin, "wb") as fh:
        np.array([n], dtype=np.int64).tofile(fh)
        np.ascontiguousarray(tau, dtype=np.float64).tofile(fh)
        np.ascontiguousarray(prev, dtype=np.float64).tofile(fh)
    subprocess.run([exe, kind, fin, fout], check=True)
    raw = np.fromfile(fout, dtype=np.uint8)
    x = raw[:8 * n * 8].view(np.float64).reshape(8, n)
    rest = raw[8 * n * 8:].view(np.int32).reshape(3, n)
    os.remove(fin), os.remove(fout)
    return x, rest[0], rest[1], rest[2]


def report(g, x, mode, it, n):
    ref_ok, ok = g['success'][:n], mode == 0
    print("success flag agreement %.4f   host ok & ref not: %s   ref ok & host not: %s" % (
        (ok == ref_ok).mean(), np.nonzero(ok & ~ref_ok)[0], np.nonzero(~ok & ref_ok)[0]))
    both = ok & ref_ok
    xr = g['x_raw'][:, :n]
    rel = (np.abs(x - xr) / np.maximum(1, np.abs(xr))).max(0)
    print("both succeed %d: literal agreement with the reference's x  <=1e-5: %.4f  <=1e-4: %.4f  <=5e-3: %.4f" % (
        both.sum(), (rel[both] <= 1e-5).mean(), (rel[both] <= 1e-4).mean(), (rel[both] <= 5e-3).mean()))
    print("iteration count equal (both succeed): %.4f;  mean iterations host %.2f ref %.2f (all demands: %.2f / %.2f)" % (
        (it[both] == g['slsqp_nit'][:n][both]).mean(), it[both].mean(), g['slsqp_nit'][:n][both].mean(), it.mean(),
        g['slsqp_nit'][:n].mean()))
    far = np.nonzero(both)[0][rel[both] > 5e-3]
    print("other end point (> 5e-3): %d %s" % (len(far), far[:30]))
    return rel


if __name__ == "__main__":
    kind = sys.argv[1] if len(sys.argv) > 1 else "double"
    g = np.load(os.path.join(ROOT, "tests", "golden", "qp_config1.npz"))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else g['tau'].shape[1]
    exe = build()
    import time
    t0 = time.time()
    x, mode, it, mask = run(exe, kind, g['tau'][:, :n], g['prev'][:, :n])
    print("%s: %d demands in %.2f s on one host core" % (kind, n, time.time() - t0))
    report(g, x, mode, it, n)
