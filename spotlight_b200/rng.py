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


_SCRATCH = {}
_JUMP = {}
_PARALLEL_MIN_BLOCKS = 2048     # >= 2 slices of 1024 blocks: one jump round (~0.35 ms) already pays


def _jump_table(dev):
    """Device copy of the jump polynomials (data/mt19937_jump.npy) + state scratch."""
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _JUMP:
        import os
        tab = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data',
                                   'mt19937_jump.npy'))
        _JUMP[key] = (torch.from_numpy(tab.view(np.int32)).to(dev), int(tab.shape[0]),
                      torch.empty(128 * _N, dtype=torch.int32, device=dev))
    return _JUMP[key]


def _scratch(dev, nwords, ws_bytes):
    """Persistent (blocks, workspace, cursor) per device and stream: the sampler
    runs every epoch, and a fresh cudaMalloc would synchronise the device."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    cur = _SCRATCH.get(key)
    if cur is None or cur[0].numel() < nwords or cur[1].numel() < ws_bytes:
        grow = 1.5 if cur is not None else 1.0
        cur = (torch.empty(int(nwords * grow), dtype=torch.int32, device=dev),
               torch.empty(int(ws_bytes * grow) + 4096, dtype=torch.uint8, device=dev),
               torch.empty(2, dtype=torch.int64, device=dev),
               torch.empty(_N, dtype=torch.int32).pin_memory(),
               torch.empty(2, dtype=torch.int64).pin_memory())
        _SCRATCH[key] = cur
    return cur


def reserve(num_items, count, device):
    """Size the persistent sampler scratch for draws of up to ``count`` values."""
    rng = int(num_items) - 1
    if rng <= 0 or count <= 0:
        return
    lib = _lib.load()
    p_accept = (rng + 1) / float(_mask_for(rng) + 1)
    want = min(int(count), _MAX_CHUNK)
    need_words = want / p_accept + 8.0 * math.sqrt(want * (1 - p_accept)) / p_accept + 64
    nwords = (int(math.ceil((_N + need_words) / _N)) + 1) * _N
    _scratch(torch.device(device), nwords, lib.slb_sample_workspace_bytes(nwords))


def sample_items_device(num_items, shape, random_state, device, out=None):
    """``random_state.randint(0, num_items, shape, dtype=int64)`` as a CUDA tensor.

    Advances ``random_state`` exactly as the NumPy call would.  ``out``: optional
    preallocated int64 CUDA tensor with ``prod(shape)`` elements.
    """
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
    count = int(np.prod(shape)) if len(shape) else 1
    dev = torch.device(device)
    if out is None:
        out = torch.empty(count, dtype=torch.int64, device=dev)
    else:
        out = out.reshape(-1)
        if out.numel() != count or out.dtype != torch.int64 or not out.is_contiguous():
            raise ValueError('sample_items_device: out must be a contiguous int64 tensor of %d' % count)
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
        nwords = nblocks * _N
        blocks, ws, cursor, pin_key, pin_cur = _scratch(dev, nwords, lib.slb_sample_workspace_bytes(nwords))
        pin_key.copy_(torch.from_numpy(key.view(np.int32)))
        blocks[:_N].copy_(pin_key, non_blocking=True)
        pin_cur[0], pin_cur[1] = pos, 0
        cursor.copy_(pin_cur, non_blocking=True)
        if nblocks >= _PARALLEL_MIN_BLOCKS:
            table, rows, states = _jump_table(dev)
            _lib.check(lib.slb_mt19937_fill_parallel(_ptr(blocks), nblocks, _ptr(table), rows,
                                                     _ptr(states), _stream()), 'mt19937_fill_parallel')
        else:
            _lib.check(lib.slb_mt19937_fill(_ptr(blocks), nblocks, _stream()), 'mt19937_fill')
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
