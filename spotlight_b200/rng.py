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


_GAMMA = 0.5772156649015329


def _harmonic(x):
    if x < 64:
        return sum(1.0 / k for k in range(1, int(x) + 1))
    return math.log(x) + _GAMMA + 0.5 / x - 1.0 / (12.0 * x * x)


def _shuffle_expected_words(n):
    """E[stream words consumed by RandomState.shuffle of n elements] = sum_i M(i)/(i+1)."""
    total, hi = 0.0, n - 1
    while hi >= 1:
        M = _mask_for(hi) + 1
        lo = M // 2
        total += M * (_harmonic(hi + 1) - _harmonic(lo))
        hi = lo - 1
    return total


_SHUFFLE_WS = {}
SHUFFLE_DEVICE_MAX = 1 << 29


def shuffled_order_device(n, random_state, device, rounds=40):
    """``random_state.shuffle(arange(n))`` as an int64 CUDA tensor, computed on the device.

    Bit-exact with NumPy (same permutation, ``random_state`` left in the same state):
    the stream is generated by the jump-ahead MT19937 kernels, the acceptance pattern of
    the n-1 masked-rejection draws and the chain of dependent swaps are both resolved in
    parallel (csrc/shuffle.cu).  Replaces spotlight/torch_utils.py:46-47 for the epoch
    shuffle of factorization/implicit.py:212-214.
    """
    return shuffle_end(shuffle_begin(n, random_state, device, rounds))


def shuffle_begin(n, random_state, device, rounds=40, margin=8.0):
    """Launches the device shuffle (asynchronous) and returns a handle for
    :func:`shuffle_end`; ``random_state`` is only read here."""
    n = int(n)
    dev = torch.device(device)
    h = dict(n=n, dev=dev, rs=random_state, rounds=int(rounds), margin=float(margin))
    if n <= 1:
        return h
    if n > SHUFFLE_DEVICE_MAX:
        raise ValueError('shuffled_order_device: n must be <= 2**29')
    lib = _lib.load()
    st = random_state.get_state()
    if st[0] != 'MT19937':
        raise ValueError('shuffled_order_device needs a legacy MT19937 RandomState')
    key = np.ascontiguousarray(st[1], dtype=np.uint32)
    pos = int(st[2])
    need_words = _shuffle_expected_words(n) + margin * math.sqrt(2.0 * n) + 64
    nblocks = int(math.ceil((pos + need_words) / _N)) + 1
    nwords = nblocks * _N
    ws_bytes = lib.slb_shuffle_workspace_bytes(n, nwords - pos)
    blocks, _, _, pin_key, _ = _scratch(dev, nwords, 0)
    skey = (dev.index if dev.index is not None else torch.cuda.current_device(),
            torch.cuda.current_stream(dev).cuda_stream)
    cur = _SHUFFLE_WS.get(skey)
    if cur is None or cur[0].numel() < ws_bytes:
        cur = (torch.empty(ws_bytes + 4096, dtype=torch.uint8, device=dev),
               torch.empty(4, dtype=torch.int64, device=dev))
        _SHUFFLE_WS[skey] = cur
    ws, cursor = cur
    pin_key.copy_(torch.from_numpy(key.view(np.int32)))
    blocks[:_N].copy_(pin_key, non_blocking=True)
    if nblocks >= _PARALLEL_MIN_BLOCKS:
        table, rows, states = _jump_table(dev)
        _lib.check(lib.slb_mt19937_fill_parallel(_ptr(blocks), nblocks, _ptr(table), rows,
                                                 _ptr(states), _stream()), 'mt19937_fill_parallel')
    else:
        _lib.check(lib.slb_mt19937_fill(_ptr(blocks), nblocks, _stream()), 'mt19937_fill')
    order = torch.empty(n, dtype=torch.int64, device=dev)
    h.update(st=st, pos=pos, nwords=nwords, blocks=blocks, ws=ws, cursor=cursor, order=order)
    _shuffle_launch(h, 0)
    return h


def _shuffle_launch(h, resume):
    lib = _lib.load()
    rc = lib.slb_shuffle_order(_ptr(h['blocks']), h['nwords'], _ptr(h['cursor']), h['pos'], h['n'],
                               h['rounds'], resume, _ptr(h['order']), _ptr(h['ws']), h['ws'].numel(),
                               _stream())
    _lib.check(rc, 'shuffle_order')


def shuffle_end(h):
    """Waits for the permutation, hands the generator state back to ``random_state``."""
    n = h['n']
    if n <= 1:
        return torch.zeros(n, dtype=torch.int64, device=h['dev'])
    while True:
        end, swaps, converged, _ = (int(v) for v in h['cursor'].tolist())     # the call's one sync
        if not converged:
            _shuffle_launch(h, 1)               # more global rounds from the workspace's state
            continue
        if swaps == n - 1:
            break
        # stream too short (an 8-sigma event): regenerate a longer one
        h = shuffle_begin(n, h['rs'], h['dev'], h['rounds'], h['margin'] * 4.0)
    st, blocks = h['st'], h['blocks']
    if end % _N == 0 and end > 0:           # numpy leaves pos = 624 on a block boundary
        blk, pos = end // _N - 1, _N
    else:
        blk, pos = end // _N, end % _N
    key = blocks[blk * _N:(blk + 1) * _N].cpu().numpy().view(np.uint32).copy()
    h['rs'].set_state(('MT19937', key, pos, st[3], st[4]))
    return h['order']


def permute_ids(order, users, items=None):
    """``(users[order], items[order])`` as int64 CUDA tensors in one pass (the gathers of
    spotlight/torch_utils.py:49-52); ``users`` / ``items`` are int32 or int64 CUDA tensors."""
    lib = _lib.load()
    n = order.numel()
    if order.dtype != torch.int64 or not order.is_cuda:
        raise ValueError('permute_ids: order must be an int64 CUDA tensor')
    srcs = [users] if items is None else [users, items]
    if any(t.dtype != users.dtype or t.numel() != n or not t.is_cuda for t in srcs) or \
            users.dtype not in (torch.int32, torch.int64):
        raise ValueError('permute_ids: ids must be CUDA int32/int64 tensors as long as order')
    srcs = [t.contiguous() for t in srcs]
    outs = [torch.empty(n, dtype=torch.int64, device=order.device) for _ in srcs]
    rc = lib.slb_permute_ids(_ptr(order.contiguous()), n, _ptr(srcs[0]), _ptr(srcs[1]) if items is not None else None,
                             users.element_size(), _ptr(outs[0]), _ptr(outs[1]) if items is not None else None,
                             _stream())
    _lib.check(rc, 'permute_ids')
    return (outs[0], outs[1]) if items is not None else outs[0]
