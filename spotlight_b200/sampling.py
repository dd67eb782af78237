"""Negative sampling (reference: spotlight/sampling.py:8-36).

``sample_items`` keeps the reference signature and semantics: uniform ids from
``random_state.randint(0, num_items, shape, dtype=int64)`` with no rejection
of positives.  Pass ``device=`` to draw on the GPU: the result is bit-identical
to the NumPy call and ``random_state`` is advanced identically
(spotlight_b200/rng.py).
"""

import numpy as np


def sample_items(num_items, shape, random_state=None, device=None, out=None):
    """Randomly sample item ids in ``[0, num_items)``.

    Returns a NumPy int64 array (``device is None``, host path identical to the
    reference) or a CUDA int64 tensor drawn by the device MT19937 stream.
    """
    if random_state is None:
        random_state = np.random.RandomState()
    if device is None:
        return random_state.randint(0, num_items, shape, dtype=np.int64)
    from spotlight_b200.rng import sample_items_device
    return sample_items_device(num_items, shape, random_state, device, out=out)
