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
_STATES = {}
_PARALLEL_MIN_BLOCKS = 512      # >= 2 slices of 256 blocks: one jump round (~0.35 ms) already pays
_J0_LOG2 = 8                    # fine stride of the direct jump table (data/gen_mt19937_jump.py)
_SIGMA = 12.0                   # stream-length margin of a device draw (chained draws cannot re-draw)


def _dev_key(dev):
    return dev.index if dev.index is not None else torch.cuda.current_device()


def _jump_table(dev):
    """Device copies of the jump polynomials (data/mt19937_jump*.npy)."""
    key = _dev_key(dev)
    if key not in _JUMP:
        import os
        here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
        tab = np.load(os.path.join(here, 'mt19937_jump.npy'))
        direct = np.load(os.path.join(here, 'mt19937_jump_direct.npy'))
        _JUMP[key] = (torch.from_numpy(tab.view(np.int32)).to(dev), int(tab.shape[0]),
                      torch.from_numpy(direct.view(np.int32)).to(dev), int(direct.shape[0]))
    return _JUMP[key]


def _states(dev, slots):
    """Jump-state scratch per (device, stream): the shuffle and the sampler may run on
    different streams at the same time."""
    key = (_dev_key(dev), torch.cuda.current_stream(dev).cuda_stream)
    cur = _STATES.get(key)
    if cur is None or cur.numel() < slots * _N:
        cur = torch.empty(max(int(slots * 1.5), 512) * _N, dtype=torch.int32, device=dev)
        _STATES[key] = cur
    return cur


def generate_blocks(blocks, nblocks, dev):
    """Fill blocks[1..nblocks) from blocks[0] on the current stream (asynchronous)."""
    lib = _lib.load()
    if nblocks >= _PARALLEL_MIN_BLOCKS:
        table, rows, direct, drows = _jump_table(dev)
        slots = int(lib.slb_mt19937_direct_slots(nblocks, _J0_LOG2))
        states = _states(dev, slots)
        _lib.check(lib.slb_mt19937_fill_direct(_ptr(blocks), nblocks, _ptr(table), rows, _ptr(direct),
                                               drows, _J0_LOG2, _ptr(states), states.numel() // _N,
                                               _stream()), 'mt19937_fill_direct')
    else:
        _lib.check(lib.slb_mt19937_fill(_ptr(blocks), nblocks, _stream()), 'mt19937_fill')


def _scratch(dev, nwords, ws_bytes):
    """Persistent (blocks, workspace, cursor) per device and stream: the sampler
    runs every epoch, and a fresh cudaMalloc would synchronise the device."""
    key = (_dev_key(dev), torch.cuda.current_stream(dev).cuda_stream)
    cur = _SCRATCH.get(key)
    if cur is None or cur[0].numel() < nwords or cur[1].numel() < ws_bytes:
        grow = 1.5 if cur is not None else 1.0
        # the small page-locked hand-over buffers are made once per (device, stream): cudaHostAlloc
        # costs milliseconds and must not recur when the device scratch grows mid-run
        pins = cur[3:] if cur is not None else (torch.empty(_N, dtype=torch.int32).pin_memory(),
                                                torch.empty(4, dtype=torch.int64).pin_memory())
        cur = (torch.empty(int(nwords * grow), dtype=torch.int32, device=dev),
               torch.empty(int(ws_bytes * grow) + 4096, dtype=torch.uint8, device=dev),
               torch.empty(4, dtype=torch.int64, device=dev)) + tuple(pins)
        _SCRATCH[key] = cur
    return cur


def _words_for(count, p_accept, pos=0):
    """Stream words (whole blocks) that hold `count` accepted values with a _SIGMA margin."""
    need = count / p_accept + _SIGMA * math.sqrt(count * (1 - p_accept)) / p_accept + 64
    return (int(math.ceil((pos + need) / _N)) + 1) * _N


def reserve(num_items, count, device):
    """Size the persistent sampler scratch for draws of up to ``count`` values."""
    rng = int(num_items) - 1
    if rng <= 0 or count <= 0:
        return
    lib = _lib.load()
    p_accept = (rng + 1) / float(_mask_for(rng) + 1)
    nwords = _words_for(min(int(count), _MAX_CHUNK), p_accept, _N)
    dev = torch.device(device)
    _scratch(dev, nwords, lib.slb_sample_workspace_bytes(nwords))
    if nwords // _N >= _PARALLEL_MIN_BLOCKS:
        _jump_table(dev)
        _states(dev, int(lib.slb_mt19937_direct_slots(nwords // _N, _J0_LOG2)))


class DeviceStream(object):
    """The model's ``RandomState`` taken over by the device for a run of draws.

    ``draw`` enqueues a bit-exact ``randint(0, num_items, count)`` and leaves the
    generator hand-over on the device (``slb_sample_bounded_chain``), so any number
    of draws -- one per minibatch in the reference, implicit.py:256-259 -- chain on
    the current stream without a host round trip.  ``finish`` is the single
    synchronisation: it reads the (key, pos) pair back and ``set_state``s it on the
    host ``RandomState``, which is then exactly where NumPy would have left it.
    All calls must be made with the same current stream.
    """

    def __init__(self, random_state, device):
        self.rs, self.dev = random_state, torch.device(device)
        st = random_state.get_state()
        if st[0] != 'MT19937':
            raise ValueError('DeviceStream needs a legacy MT19937 RandomState')
        self._gauss = (st[3], st[4])
        self._blocks = None
        self._key = np.ascontiguousarray(st[1], dtype=np.uint32)
        self._pos = int(st[2])
        self._open = False
        self._drawn = 0

    def _ensure(self, nwords, ws_bytes):
        blocks, ws, cursor, pin_key, pin_cur = _scratch(self.dev, nwords, ws_bytes)
        if not self._open:
            pin_key.copy_(torch.from_numpy(self._key.view(np.int32)))
            blocks[:_N].copy_(pin_key, non_blocking=True)
            pin_cur[0], pin_cur[1], pin_cur[2], pin_cur[3] = self._pos, 0, 0, 0
            cursor.copy_(pin_cur, non_blocking=True)
            self._open = True
        elif blocks is not self._blocks:            # scratch grew: carry the device state over
            blocks[:_N].copy_(self._blocks[:_N])
            cursor.copy_(self._cursor)
        self._blocks, self._ws, self._cursor, self._pin_key, self._pin_cur = blocks, ws, cursor, pin_key, pin_cur

    def draw(self, num_items, count, out=None):
        """``randint(0, num_items, count, dtype=int64)`` into ``out`` (asynchronous)."""
        count = int(count)
        if out is None:
            out = torch.empty(count, dtype=torch.int64, device=self.dev)
        out = out.reshape(-1)
        if out.numel() != count or out.dtype != torch.int64 or not out.is_contiguous():
            raise ValueError('DeviceStream.draw: out must be a contiguous int64 tensor of %d' % count)
        rng = int(num_items) - 1
        if rng < 0:
            raise ValueError('num_items must be positive')
        if count == 0:
            return out
        if rng == 0:                       # numpy consumes no randomness for a 1-value range
            return out.zero_()
        if rng >= 0xFFFFFFFF:
            raise ValueError('num_items must be < 2**32')
        lib = _lib.load()
        p_accept = (rng + 1) / float(_mask_for(rng) + 1)
        done = 0
        while done < count:
            want = min(count - done, _MAX_CHUNK)
            nwords = _words_for(want, p_accept, _N)       # the hand-over position is <= 624
            self._ensure(nwords, lib.slb_sample_workspace_bytes(nwords))
            generate_blocks(self._blocks, nwords // _N, self.dev)
            rc = lib.slb_sample_bounded_chain(_ptr(self._blocks), nwords, _ptr(self._cursor),
                                              ctypes.c_uint32(rng), want, _ptr(out[done:done + want]),
                                              _ptr(self._ws), self._ws.numel(), _stream())
            _lib.check(rc, 'sample_bounded_chain')
            done += want
            self._drawn += want
        return out

    def finish(self):
        """Synchronise, hand the generator back to the host ``RandomState``."""
        if not self._open:
            return
        self._pin_key.copy_(self._blocks[:_N], non_blocking=True)
        self._pin_cur.copy_(self._cursor, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        pos, _, short, produced = (int(v) for v in self._pin_cur.tolist())
        key = self._pin_key.numpy().view(np.uint32).copy()
        self._open = False
        if short or produced != self._drawn:
            raise RuntimeError('device sampler ran out of stream words (a %g-sigma event): %d of %d '
                               'values drawn' % (_SIGMA, produced, self._drawn))
        self.rs.set_state(('MT19937', key, pos, self._gauss[0], self._gauss[1]))
        self._key, self._pos, self._drawn = key, pos, 0


def sample_items_device(num_items, shape, random_state, device, out=None):
    """``random_state.randint(0, num_items, shape, dtype=int64)`` as a CUDA tensor.

    Advances ``random_state`` exactly as the NumPy call would.  ``out``: optional
    preallocated int64 CUDA tensor with ``prod(shape)`` elements.
    """
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(s) for s in shape)
    count = int(np.prod(shape)) if len(shape) else 1
    stream = DeviceStream(random_state, device)
    res = stream.draw(num_items, count, out)
    stream.finish()
    return res.reshape(shape)


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
    generate_blocks(blocks, nblocks, dev)
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
