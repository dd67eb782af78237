"""Prediction id handling (reference: spotlight/factorization/_components.py:8-25)."""

import numpy as np
import torch

from spotlight_b200.torch_utils import gpu


def _predict_process_ids(user_ids, item_ids, num_items, use_cuda):
    """Broadcast a scalar user over ``item_ids`` (all items when None) and
    return flat int64 tensors on the model's device."""
    if item_ids is None:
        item_ids = np.arange(num_items, dtype=np.int64)
    if np.isscalar(user_ids):
        user_ids = np.array(user_ids, dtype=np.int64)
    users = torch.from_numpy(np.asarray(user_ids).reshape(-1).astype(np.int64))
    items = torch.from_numpy(np.asarray(item_ids).reshape(-1).astype(np.int64))
    if items.size(0) != users.size(0):
        users = users.expand(items.size(0)).contiguous()
    return gpu(users, use_cuda), gpu(items, use_cuda)
