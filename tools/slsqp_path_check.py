"""Compare tools/slsqp_path_proto.py with SciPy's SLSQP on the config-1 golden batch (build container, CPU).
    python tools/slsqp_path_check.py [n]
Prints: exit-mode agreement, distance of the end points, and for a few demands the per-iteration distance of the iterates
(SciPy's are captured through its `callback`)."""
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import slsqp_path_proto as SP  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "qp_config1.npz"))
tau, prev = g['tau'], g['prev']


def work(j):
    x, mode, it = SP.slsqp(tau[:, j], prev[:, j])
    return x, mode, it


def scipy_trace(j):
    from scipy.optimize import minimize
    from oracle import qp_oracle as QO
    p = [float(v) for v in prev[:, j]]
    x0 = np.array(p + [0.0, 0.0, 0.0])
    tr = []
    minimize(lambda x: QO._objective(x, p), x0, method='SLSQP', bounds=QO._bounds(),
             constraints=QO._constraints(tau[:, j], p, analytic=False), callback=lambda xk: tr.append(np.array(xk)))
    return tr


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else tau.shape[1]
    with mp.Pool(os.cpu_count()) as pool:
        rows = pool.map(work, range(n), chunksize=32)
    X = np.array([r[0] for r in rows]).T
    MODE = np.array([r[1] for r in rows])
    IT = np.array([r[2] for r in rows])
    ref_mode, ref_it = g['slsqp_status'][:n], g['slsqp_nit'][:n]
    print("exit mode agreement %.4f   (mismatches: %s)" % ((MODE == ref_mode).mean(), np.nonzero(MODE != ref_mode)[0][:20]))
    print("success flag agreement %.4f  proto ok & ref not: %s   ref ok & proto not: %s" % (
        ((MODE == 0) == (ref_mode == 0)).mean(), np.nonzero((MODE == 0) & (ref_mode != 0))[0],
        np.nonzero((MODE != 0) & (ref_mode == 0))[0]))
    print("iteration count equal %.4f" % (IT == ref_it).mean())
    ok = (MODE == 0) & (ref_mode == 0)
    d = np.abs(X[:, ok] - g['x_raw'][:, :n][:, ok]).max(0)
    print("both succeed: %d; end-point distance pct 50/90/99/max: %s" % (ok.sum(), np.percentile(d, [50, 90, 99, 100])))
    far = np.nonzero(ok)[0][d > 5e-3]
    print("end points farther than 5e-3: %d %s" % (len(far), far[:20]))
    for j in list(far[:3]) + list(np.nonzero(MODE != ref_mode)[0][:3]):
        tr_ref, tr = scipy_trace(j), []
        SP.slsqp(tau[:, j], prev[:, j], trace=tr)
        m = min(len(tr), len(tr_ref))
        print("demand %d: scipy %d iterates (mode %d), proto %d (mode %d); per-iteration distance:" %
              (j, len(tr_ref), ref_mode[j], len(tr), MODE[j]),
              ["%.1e" % np.abs(tr[k] - tr_ref[k]).max() for k in range(m)])
