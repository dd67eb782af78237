"""CPU restatement of the *parallel* formulation of ``RandomState.shuffle(arange(n))`` that
``spotlight_b200/csrc/shuffle.cu`` implements (reference call site
spotlight/torch_utils.py:46-47).  Test infrastructure only: the product never imports it.

NumPy's loop is ``for i = n-1 .. 1: j = rk_interval(i); x[i], x[j] = x[j], x[i]`` with
``rk_interval(i)`` = the first stream word that, masked to ``bit_length(i)`` bits, is <= i.
``resolve_draws`` finds which stream words are accepted, and for which step, as the fixed
point of (acceptance flags -> accepted-before counts -> flags); ``apply_swaps`` turns the
n-1 dependent swaps into next-occurrence links and pointer chases.  Pinned against NumPy
itself in tests/test_oracle.py.
"""

import numpy as np


def _smear(i):
    m = i.astype(np.uint32).copy()
    for s in (1, 2, 4, 8, 16):
        m |= m >> np.uint32(s)
    return m


def resolve_draws(words, n):
    """words: the tempered uint32 stream from the generator's current position.
    Returns (j, used): j[i] for every step i (j[0] = 0) and the number of words consumed."""
    words = np.asarray(words, dtype=np.uint32)
    A = np.zeros(len(words), dtype=np.int64)          # accepted words before word t
    while True:
        i = n - 1 - A                                 # the step word t would serve
        live = i >= 1
        ic = np.where(live, i, 1)
        flag = live & ((words & _smear(ic)) <= ic.astype(np.uint32))
        nxt = np.cumsum(flag) - flag
        if np.array_equal(nxt, A):
            break
        A = nxt
    acc = np.nonzero(flag)[0]
    if len(acc) != n - 1:
        raise ValueError('stream too short: %d of %d draws' % (len(acc), n - 1))
    steps = n - 1 - A[acc]
    j = np.zeros(n, dtype=np.int64)
    j[steps] = words[acc] & _smear(steps)
    return j, (int(acc[-1]) + 1 if n > 1 else 0)


def apply_swaps(j):
    """The permutation the swaps (i, j[i]), i = n-1 .. 1, leave in arange(n)."""
    n = len(j)
    idx = np.arange(n, dtype=np.int64)
    order = np.lexsort((idx, j))                      # steps grouped by target, ascending
    js = j[order]
    same = js[1:] == js[:-1]
    parent = np.full(n, -1, dtype=np.int64)           # next larger step on the same target
    parent[order[:-1][same]] = order[1:][same]
    first = np.full(n, -1, dtype=np.int64)            # smallest step targeting position v
    starts = np.r_[True, ~same]
    first[js[starts]] = order[starts]
    m = np.where(first == idx, parent, first)         # smallest step > x targeting position x
    x = idx.copy()                                    # V(x): follow m until nobody targets x
    cur = m.copy()
    active = cur >= 0
    while active.any():
        x[active] = cur[active]
        cur = np.where(active, m[x], -1)
        active = cur >= 0
    return np.where(parent >= 0, x[np.where(parent >= 0, parent, 0)], j)
