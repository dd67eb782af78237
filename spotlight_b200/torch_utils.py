"""Batching / device helpers with the reference's names and behaviour
(spotlight/torch_utils.py:6-69)."""

import numpy as np
import torch


def gpu(tensor, gpu=False):
    """Move to the current CUDA device when ``gpu`` is set (torch_utils.py:6-11)."""
    return tensor.cuda() if gpu else tensor


def cpu(tensor):
    """Bring a tensor back to host memory (torch_utils.py:14-19)."""
    return tensor.cpu() if tensor.is_cuda else tensor


def minibatch(*tensors, **kwargs):
    """Contiguous slices of ``batch_size``; the last one may be short
    (torch_utils.py:22-32)."""
    batch_size = kwargs.get('batch_size', 128)
    n = len(tensors[0])
    single = len(tensors) == 1
    for lo in range(0, n, batch_size):
        hi = lo + batch_size
        yield tensors[0][lo:hi] if single else tuple(t[lo:hi] for t in tensors)


def shuffle(*arrays, **kwargs):
    """One permutation from ``random_state`` applied to every array
    (torch_utils.py:35-52).  The permutation is ``random_state.shuffle`` of
    ``arange(n)`` so the MT19937 stream advances exactly as in the reference."""
    random_state = kwargs.get('random_state')
    lengths = set(len(a) for a in arrays)
    if len(lengths) != 1:
        raise ValueError('All inputs to shuffle must have the same length.')
    if random_state is None:
        random_state = np.random.RandomState()
    order = np.arange(lengths.pop())
    random_state.shuffle(order)
    if len(arrays) == 1:
        return arrays[0][order]
    return tuple(a[order] for a in arrays)


def shuffled_order(n, random_state):
    """The permutation ``random_state.shuffle(np.arange(n))`` produces, leaving
    ``random_state`` in exactly the state NumPy would (torch_utils.py:46-47), computed
    by the library's prefetching Fisher-Yates (csrc/host_shuffle.cpp, ~3x NumPy)."""
    import ctypes
    from spotlight_b200 import _lib
    lib = _lib.load()
    st = random_state.get_state()
    if st[0] != 'MT19937' or n - 1 > 0xFFFFFFFE:
        order = np.arange(n)
        random_state.shuffle(order)
        return order
    key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
    pos = ctypes.c_int32(int(st[2]))
    order = np.empty(n, dtype=np.int32 if n < 2**31 else np.int64)
    rc = lib.slb_host_shuffle_order(key.ctypes.data, ctypes.byref(pos), n, order.itemsize,
                                    order.ctypes.data)
    _lib.check(rc, 'host_shuffle_order')
    random_state.set_state(('MT19937', key, pos.value, st[3], st[4]))
    return order


def assert_no_grad(variable):
    if variable.requires_grad:
        raise ValueError(
            "nn criterions don't compute the gradient w.r.t. targets - please "
            "mark these variables as volatile or not requiring gradients")


def set_seed(seed, cuda=False):
    """Seed torch's global generators (torch_utils.py:64-69)."""
    torch.manual_seed(seed)
    if cuda:
        torch.cuda.manual_seed(seed)
