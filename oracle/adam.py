"""Lazy-exact row-wise Adam, restated on the CPU (TEST INFRASTRUCTURE ONLY).

The reference's default optimizer is dense ``torch.optim.Adam(params, weight_decay=l2, lr)``
(spotlight/factorization/implicit.py:143-148): every row moves at every step, also rows without
a gradient (their first moment decays).  The product (spotlight_b200/csrc/mf_adam.cuh) applies it
row-wise and lazily; this module states the same scheme in NumPy float64 so that the *scheme*
-- catch-up of every referenced row BEFORE the forward pass, real step on the touched rows,
flush at the end -- can be pinned on the CPU against the reference's recorded default-Adam
trajectory (tests/golden/fit_pointwise_adam.npz) and against torch.optim.Adam.

Per element and step t (torch/optim/adam.py, _single_tensor_adam):
    g += wd * w;  m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g
    w -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
"""

import numpy as np


class LazyAdamTable(object):
    """One (rows, D) table with its Adam state and the step each row is current for."""

    def __init__(self, w, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.w = np.array(w, dtype=np.float64)
        self.m = np.zeros_like(self.w)
        self.v = np.zeros_like(self.w)
        self.last = np.zeros(self.w.shape[0], dtype=np.int64)
        self.lr, (self.b1, self.b2), self.eps, self.wd = lr, betas, eps, weight_decay

    def _step(self, rows, t, g):
        g = g + self.wd * self.w[rows]
        self.m[rows] += (g - self.m[rows]) * (1.0 - self.b1)
        self.v[rows] = self.v[rows] * self.b2 + (1.0 - self.b2) * g * g
        ss, bc = self.lr / (1.0 - self.b1 ** t), np.sqrt(1.0 - self.b2 ** t)
        self.w[rows] -= ss * (self.m[rows] / (np.sqrt(self.v[rows]) / bc + self.eps))

    def catch_up(self, rows, upto):
        """Replay the gradient-free steps (last, upto] of ``rows`` (mf_adam_prepass_kernel)."""
        rows = np.unique(np.asarray(rows))
        for r in rows:
            for t in range(int(self.last[r]) + 1, upto + 1):
                self._step(np.array([r]), t, 0.0)
            self.last[r] = max(int(self.last[r]), upto)

    def apply(self, rows, grads, t):
        """The real step t on the touched rows (mf_adam_apply_kernel); rows must be current for t-1."""
        rows = np.asarray(rows)
        assert (self.last[rows] == t - 1).all()
        self._step(rows, t, grads)
        self.last[rows] = t

    def flush(self, t):
        """Every row current for step t (adam_flush_kernel)."""
        self.catch_up(np.arange(self.w.shape[0]), t)
