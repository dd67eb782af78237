"""Row-wise optimizers fused into the training step.

The reference's default ``optim.Adam`` (spotlight/factorization/implicit.py:
143-148) sweeps every embedding row each minibatch -- 72 % of its step at
1M x 100K x 64.  SGD and Adagrad change a row only when its gradient is
non-zero, so applying them to the touched rows alone is *exactly* the dense
update (no weight decay / momentum); here that update runs in
``mf_apply_kernel`` right after the gradient kernel, never materialising a
dense gradient.

Both classes are ordinary ``torch.optim`` optimizers (usable with any module
via ``.step()``); ``ImplicitFactorizationModel.fit`` recognises them through
``fused_kind`` and switches to the on-device epoch pipeline.

Use as ``optimizer_func``::

    model = ImplicitFactorizationModel(loss='bpr', use_cuda=True,
                                       optimizer_func=fused_adagrad(lr=0.05))
"""

import torch

from spotlight_b200 import _lib


class FusedSGD(torch.optim.SGD):
    """Plain SGD (no momentum).  ``weight_decay`` is applied to touched rows only."""

    fused_kind = _lib.OPT_SGD

    def __init__(self, params, lr=1e-2, weight_decay=0.0):
        super(FusedSGD, self).__init__(params, lr=lr, momentum=0.0, weight_decay=weight_decay)

    def fused_hparams(self):
        g = self.param_groups[0]
        return dict(lr=float(g['lr']), weight_decay=float(g['weight_decay']), eps=0.0)

    def fused_state(self, param):
        return None


class FusedAdagrad(torch.optim.Adagrad):
    """Adagrad with ``lr_decay = 0`` and ``initial_accumulator_value = 0``
    (torch defaults).  State lives in ``self.state[p]['sum']`` exactly as in
    ``torch.optim.Adagrad`` so training can continue with either."""

    fused_kind = _lib.OPT_ADAGRAD

    def __init__(self, params, lr=1e-2, weight_decay=0.0, eps=1e-10):
        super(FusedAdagrad, self).__init__(params, lr=lr, lr_decay=0.0,
                                           weight_decay=weight_decay,
                                           initial_accumulator_value=0.0, eps=eps)

    def fused_hparams(self):
        g = self.param_groups[0]
        return dict(lr=float(g['lr']), weight_decay=float(g['weight_decay']), eps=float(g['eps']))

    def fused_state(self, param):
        return self.state[param]['sum']


class _Factory(object):
    """Picklable ``optimizer_func`` (models are saved whole with ``torch.save``)."""

    def __init__(self, cls, **kwargs):
        self.cls = cls
        self.kwargs = kwargs

    def __call__(self, params):
        return self.cls(params, **self.kwargs)


def fused_sgd(lr=1e-2, weight_decay=0.0):
    """``optimizer_func`` factory for :class:`FusedSGD`."""
    return _Factory(FusedSGD, lr=lr, weight_decay=weight_decay)


def fused_adagrad(lr=1e-2, weight_decay=0.0, eps=1e-10):
    """``optimizer_func`` factory for :class:`FusedAdagrad`."""
    return _Factory(FusedAdagrad, lr=lr, weight_decay=weight_decay, eps=eps)
