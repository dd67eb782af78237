"""How far do adaptive-hinge + Adagrad trajectories move under rounding-level perturbations?
CPU-only (float64 oracle).  Problem = tests/test_sharded_gpu.py::FIT.  Output of the run kept in
DESIGN.md section 6: a 1e-7 relative perturbation of the initial item table moves *every* user
row by more than 1.6e-3 (max 0.053 on a 0.32 scale) while the epoch losses move by < 2e-5."""
import sys, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import sharded_common as sc
from oracle import mf as omf
FIT = dict(seed=33, U=3000, I=800, D=32, n=300000, B=16384, n_iter=2)
rs = np.random.RandomState(8)
params, _ = sc.make_problem(6, FIT['U'], FIT['I'], FIT['D'], 8, 0)
params = tuple(p * 0.3 for p in params)
users = rs.randint(0, FIT['U'], FIT['n']).astype(np.int32); items = rs.randint(0, FIT['I'], FIT['n']).astype(np.int32)
epochs, _ = sc.reference_epochs(FIT['seed'], users, items, FIT['I'], FIT['B'], FIT['n_iter'], 4)
flat = [b for e in epochs for b in e]
def run(dtype, perturb=0.0):
    P = [p.astype(dtype) for p in params]
    if perturb:
        P[1] = (P[1] * (1 + perturb * np.random.RandomState(0).randn(*P[1].shape))).astype(dtype)
    S = [np.zeros_like(p) for p in P]
    losses = []
    for u, i, ng in flat:
        r = omf.mf_step(P[0], P[1], P[2], P[3], u, i, ng, 'adaptive_hinge', 4, dtype)
        losses.append(float(r['loss']))
        for k, g in enumerate((r['dWu'], r['dWi'], r['dbu'], r['dbi'])):
            g = g.astype(dtype)
            S[k] += g * g
            P[k] -= (0.05 * g / (np.sqrt(S[k]) + dtype(1e-10))).astype(dtype)
    return P, losses
a, la = run(np.float64)
b, lb = run(np.float32)
c, lc = run(np.float64, 1e-7)
for nm, (x, lx) in (('f32 vs f64', (b, lb)), ('f64 perturbed 1e-7 vs f64', (c, lc))):
    err = np.abs(x[0].astype(np.float64) - a[0]).max(axis=1)
    print(nm, 'Wu scale %.3f maxerr %.3e frac rows > 1.6e-3: %.4f' % (np.abs(a[0]).max(), err.max(), (err > 1.6e-3).mean()),
          'epoch loss diff', abs(np.mean(lx[:19]) - np.mean(la[:19])), abs(np.mean(lx[19:]) - np.mean(la[19:])))
