"""Implicit-feedback loss functions with the reference's signatures
(spotlight/losses.py:18-166), computed by one fused CUDA kernel each way.

All take ``(positive_predictions, negative_predictions, mask=None)`` and return
a 0-dim tensor: the mean of the per-element loss, or ``sum(loss * mask) /
mask.sum()`` when a mask is given.
"""

from spotlight_b200 import ops


def pointwise_loss(positive_predictions, negative_predictions, mask=None):
    """Logistic loss: ``(1 - sigmoid(pos)) + sigmoid(neg)`` (losses.py:40-50)."""
    return ops.loss_op('pointwise', positive_predictions, negative_predictions, mask)


def bpr_loss(positive_predictions, negative_predictions, mask=None):
    """BPR: ``1 - sigmoid(pos - neg)`` (losses.py:82-90)."""
    return ops.loss_op('bpr', positive_predictions, negative_predictions, mask)


def hinge_loss(positive_predictions, negative_predictions, mask=None):
    """Hinge: ``max(neg - pos + 1, 0)`` (losses.py:115-124)."""
    return ops.loss_op('hinge', positive_predictions, negative_predictions, mask)


def adaptive_hinge_loss(positive_predictions, negative_predictions, mask=None):
    """Hinge against the highest of several negatives, ``negative_predictions``
    stacked on dim 0 (losses.py:164-166)."""
    return ops.loss_op('adaptive_hinge', positive_predictions, negative_predictions, mask)
