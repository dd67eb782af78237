"""Device-resident NumPy-legacy RandomState stream.

Replaces the host-side ``random_state.randint(0, num_items, shape)`` of
spotlight/sampling.py:34 (one call + one H2D copy per minibatch in the
reference, spotlight/factorization/implicit.py:256-260) with an on-device
MT19937 + masked-rejection draw that is bit-exact with NumPy, so a whole
epoch's negatives are produced by a handful of kernel launches.

The model's ``numpy.random.RandomState`` stays the single source of truth:
``sample`` takes the stream over from it and hands the exact post-draw state
back (``set_state``), because ``shuffle`` draws from the *same* stream between
epochs (spotlight/factorization/implicit.py:212-214).
"""

import ctypes
import math

import numpy as np
import torch

from spotlight_b200 import _lib
from spotlight_b200.ops import _ptr, _stream

_N = 624
_MAX_CHUNK = 1 << 26          # values per device call (bounds scratch memory)


def _mask_for(r):
    m = r
    for s in (1, 2, 4, 8, 16):
        m |= m >> s
    return m


def sample_items_device(num_items, shape, random_state, device):
    """``random_state.randint(0, num_items, shape, dtype=int64)`` as a CUDA tensor.

    Advances ``random_state`` exactly as the NumPy call would.
    """
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
    count = int(np.prod(shape)) if len(shape) else 1
    dev = torch.device(device)
    out = torch.empty(count, dtype=torch.int64, device=dev)
    rng = int(num_items) - 1
    if rng < 0:
        raise ValueError('num_items must be positive')
    if count == 0:
        return out.reshape(shape)
    if rng == 0:                       # numpy consumes no randomness for a 1-value range
        return out.zero_().reshape(shape)
    if rng >= 0xFFFFFFFF:
        raise ValueError('num_items must be < 2**32')
    lib = _lib.load()
    p_accept = (rng + 1) / float(_mask_for(rng) + 1)

    st = random_state.get_state()
    key = np.ascontiguousarray(st[1], dtype=np.uint32)
    pos = int(st[2])
    done = 0
    while done < count:
        want = min(count - done, _MAX_CHUNK)
        # words needed ~ want / p  (+ 8 sigma), plus the unread tail of block 0
        need_words = want / p_accept + 8.0 * math.sqrt(want * (1 - p_accept)) / p_accept + 64
        nblocks = int(math.ceil((pos + need_words) / _N)) + 1
        blocks = torch.empty(nblocks * _N, dtype=torch.int32, device=dev)
        blocks[:_N].copy_(torch.from_numpy(key.view(np.int32)))
        _lib.check(lib.slb_mt19937_fill(_ptr(blocks), nblocks, _stream()), 'mt19937_fill')
        cursor = torch.tensor([pos, 0], dtype=torch.int64, device=dev)
        nwords = nblocks * _N
        ws = torch.empty(lib.slb_sample_workspace_bytes(nwords), dtype=torch.uint8, device=dev)
        chunk = out[done:done + want]
        rc = lib.slb_sample_bounded(_ptr(blocks), nwords, _ptr(cursor), ctypes.c_uint32(rng), want,
                                    _ptr(chunk), _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, 'sample_bounded')
        end, produced = (int(v) for v in cursor.tolist())       # sync: once per chunk
        done += produced
        # hand the state over: block containing the next unread word
        if end >= nwords:                       # stream exhausted (rare: 8-sigma margin)
            blk, pos = nblocks - 1, _N
        elif end % _N == 0 and end > 0:         # numpy leaves pos = 624 on a block boundary
            blk, pos = end // _N - 1, _N
        else:
            blk, pos = end // _N, end % _N
        key = blocks[blk * _N:(blk + 1) * _N].cpu().numpy().view(np.uint32).copy()
    random_state.set_state(('MT19937', key, pos, 0, 0.0))
    return out.reshape(shape)
