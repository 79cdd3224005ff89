"""Row-by-row parity of an allocator implementation against the reference's outputs in tests/golden/qp_config1.npz
(QPTA.solve_QP on the 4096-demand config-1 batch, container SciPy; tests/golden/gen_golden.py).

Every demand lands in exactly one class:
  FLAG   success flag differs from the reference's                                   -> must be empty
  FAIL   both report failure (the caller holds the previous state, :267-269)         -> nothing else to compare
  A      literal: |x - x_ref| <= 1e-5 max(1, |x_ref|) on the reference's own (raw) output
  B      not literal, but equal (1e-5) to the reference's call with analytic instead of finite-difference derivatives:
         the forward differences (step 1.49e-8) tipped one of the reference's |f - f0| < 1e-6 stopping tests
  C      neither, but in the reference's basin (<= 1e-3), with the reference's active set, on a row where the reference
         does not reproduce ITSELF to 1e-5: it moves by more than that when tau is perturbed by 1e-9 (relative) or when its
         derivatives are made exact
  X      anything else = unexplained                                                 -> must be empty
"""
import numpy as np

TOL = 1e-5


def _rel(a, b):
    return (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max(axis=0)


def classify(g, x_raw, ok, mask):
    """x_raw [8, n] (before the |x| < 0.01 clean-up), ok [n] bool, mask [n] active-set bits.  Returns (cls [n] of str,
    dict of per-row distances)."""
    ref = g['success'].astype(bool)
    ok = np.asarray(ok).astype(bool)
    d_ref, d_exact = _rel(x_raw, g['x_raw']), _rel(x_raw, g['x_raw_exact'])
    self_move = np.maximum(_rel(g['x_raw_pert'], g['x_raw']), _rel(g['x_raw_exact'], g['x_raw']))
    n = len(ok)
    cls = np.full(n, 'X', dtype=object)
    cls[ok != ref] = 'FLAG'
    cls[~ok & ~ref] = 'FAIL'
    both = ok & ref
    a = both & (d_ref <= TOL)
    b = both & ~a & (d_exact <= TOL) & g['success_exact'].astype(bool)
    c = both & ~a & ~b & (d_ref <= 1e-3) & (np.asarray(mask) == g['mask_raw']) & (self_move > TOL)
    cls[a], cls[b], cls[c] = 'A', 'B', 'C'
    return cls, {'d_ref': d_ref, 'd_exact': d_exact, 'self_move': self_move}


def table(g, cls, dist, it=None, title=""):
    """Markdown summary + the itemised B / C / X / FLAG rows."""
    n = len(cls)
    lines = ["### %s" % title if title else "", "",
             "| class | rows | meaning |", "|---|---|---|"]
    meaning = {'A': "literal: within 1e-5 of the reference's own output", 'B': "equals the reference run with exact derivatives",
               'C': "reference's basin and active set; the reference itself moves > 1e-5 under a 1e-9 perturbation / exact derivatives",
               'FAIL': "both report failure (hold previous)", 'FLAG': "success flag differs", 'X': "unexplained"}
    for k in ('A', 'B', 'C', 'FAIL', 'FLAG', 'X'):
        lines.append("| %s | %d | %s |" % (k, int((cls == k).sum()), meaning[k]))
    lines += ["", "| row | class | dist. to reference | dist. to exact-derivative reference | reference's own movement | iterations (ours / ref / ref exact) |",
              "|---|---|---|---|---|---|"]
    for j in np.nonzero(np.isin(cls, ['B', 'C', 'X', 'FLAG']))[0]:
        lines.append("| %d | %s | %.1e | %.1e | %.1e | %s / %d / %d |" % (
            j, cls[j], dist['d_ref'][j], dist['d_exact'][j], dist['self_move'][j], "-" if it is None else int(it[j]),
            int(g['slsqp_nit'][j]), int(g['nit_exact'][j])))
    return "\n".join(lines)
