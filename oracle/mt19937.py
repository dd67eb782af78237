"""NumPy-legacy RandomState stream, restated (oracle; test infrastructure only).

The reference draws negatives with ``random_state.randint(0, num_items, shape,
dtype=np.int64)`` (spotlight/sampling.py:34) and shuffles with
``random_state.shuffle(np.arange(n))`` (spotlight/torch_utils.py:46-47).  Both
consume one MT19937 stream through NumPy's *legacy* masked-rejection sampler.
NumPy is a third-party dependency of the reference and is unpinned
(setup.py:13 lists only torch); legacy ``RandomState`` streams are frozen by
NumPy policy (NEP 19).  This module restates the published algorithm
(Matsumoto & Nishimura 1998 for MT19937; numpy/random/src/distributions
``buffered_bounded_masked_uint32`` / ``random_interval`` for the bounded draw)
and tests/test_oracle_rng.py pins it against ``numpy.random.RandomState``.
"""

import numpy as np

N = 624
M = 397
_UPPER = np.uint32(0x80000000)
_LOWER = np.uint32(0x7FFFFFFF)
_MATRIX_A = np.uint32(0x9908B0DF)


def seed_key(seed):
    """init_genrand: the key RandomState(seed) starts from (pos = 624)."""
    key = np.empty(N, dtype=np.uint32)
    s = int(seed) & 0xFFFFFFFF
    for i in range(N):
        key[i] = s
        s = (1812433253 * (s ^ (s >> 30)) + i + 1) & 0xFFFFFFFF
    return key


def _mix(a, b, c):
    y = (a & _UPPER) | (b & _LOWER)
    return c ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), _MATRIX_A, np.uint32(0))


def twist(key):
    """One full MT19937 state transition (624 words), vectorised in 3 rounds.

    new[k] = old/new[k+397 mod 624] ^ f(old[k], old/new[k+1]); elements
    0..226 depend only on the old block, 227..453 on round 1, 454..623 on
    round 2 (and new[0] for the last word).
    """
    old = key
    new = np.empty(N, dtype=np.uint32)
    new[0:227] = _mix(old[0:227], old[1:228], old[397:624])
    new[227:454] = _mix(old[227:454], old[228:455], new[0:227])
    new[454:623] = _mix(old[454:623], old[455:624], new[227:396])
    new[623] = _mix(old[623:624], new[0:1], new[396:397])[0]
    return new


def temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
    y = y ^ (y >> np.uint32(18))
    return y


def gen_mask(r):
    """Smallest 2^m - 1 >= r."""
    r = int(r)
    m = r
    m |= m >> 1
    m |= m >> 2
    m |= m >> 4
    m |= m >> 8
    m |= m >> 16
    m |= m >> 32
    return m


class MT19937(object):
    """Stream with the same (key, pos) state as RandomState.get_state()[1:3]."""

    def __init__(self, seed=None, state=None):
        if state is not None:
            self.key = np.array(state[0], dtype=np.uint32).copy()
            self.pos = int(state[1])
        else:
            self.key = seed_key(seed)
            self.pos = N

    @classmethod
    def from_random_state(cls, rs):
        st = rs.get_state()
        return cls(state=(st[1], st[2]))

    def to_random_state(self, rs):
        rs.set_state(('MT19937', self.key.copy(), int(self.pos), 0, 0.0))
        return rs

    def raw(self, count):
        """The next ``count`` tempered uint32 outputs."""
        out = np.empty(count, dtype=np.uint32)
        done = 0
        while done < count:
            if self.pos >= N:
                self.key = twist(self.key)
                self.pos = 0
            take = min(count - done, N - self.pos)
            out[done:done + take] = temper(self.key[self.pos:self.pos + take])
            self.pos += take
            done += take
        return out

    def _unraw(self, count):
        """Push back the last ``count`` words of the current block."""
        assert count <= self.pos
        self.pos -= count

    def bounded(self, r, count):
        """``count`` draws of the legacy masked-rejection sampler on [0, r].

        r == 0 consumes nothing (numpy returns ``off`` directly).
        """
        r = int(r)
        out = np.empty(count, dtype=np.int64)
        if r == 0:
            out[:] = 0
            return out
        assert r < 0xFFFFFFFF
        mask = np.uint32(gen_mask(r))
        done = 0
        while done < count:
            need = count - done
            # draw roughly what is needed, then push back what was not used
            chunk = max(64, int(need * (int(mask) + 1) / (r + 1) * 1.05) + 16)
            saved_key, saved_pos = self.key.copy(), self.pos
            words = self.raw(chunk) & mask
            ok = np.nonzero(words <= np.uint32(r))[0]
            if len(ok) >= need:
                used = int(ok[need - 1]) + 1
                out[done:] = words[ok[:need]]
                # rewind to exactly ``used`` words consumed
                self.key, self.pos = saved_key, saved_pos
                self.raw(used)
                done = count
            else:
                out[done:done + len(ok)] = words[ok]
                done += len(ok)
        return out

    def randint(self, low, high, size):
        """RandomState.randint(low, high, size, dtype=int64), C order.

        spotlight/sampling.py:34 (low=0) and the ctor draw
        spotlight/factorization/implicit.py:114.
        """
        shape = (size,) if np.isscalar(size) else tuple(size)
        count = int(np.prod(shape)) if shape else 1
        vals = self.bounded(int(high) - 1 - int(low), count) + int(low)
        return vals.reshape(shape)

    def shuffle_indices(self, n):
        """``random_state.shuffle(np.arange(n))`` (torch_utils.py:46-47).

        Fisher-Yates from the end: for i = n-1 .. 1: j = bounded(i); swap.
        Pure-Python loop: small n only.
        """
        x = np.arange(n)
        for i in range(n - 1, 0, -1):
            j = int(self.bounded(i, 1)[0])
            x[i], x[j] = x[j], x[i]
        return x


def untemper(y):
    """Inverse of :func:`temper` (used to recover a key block from outputs)."""
    y = np.asarray(y, dtype=np.uint32).copy()
    y ^= y >> np.uint32(18)
    y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
    # invert y ^= (y << 7) & 0x9D2C5680 : 7 bits at a time
    t = y.copy()
    for _ in range(5):
        t = y ^ ((t << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = t
    # invert y ^= y >> 11
    t = y.copy()
    for _ in range(3):
        t = y ^ (t >> np.uint32(11))
    return t


def jump(key, poly_words):
    """State ``g(T) key`` for the jump polynomial ``g`` (624 uint32 words, bit i of
    g = bit i%32 of word i//32), T = the one-word MT19937 transition.  Horner,
    one coefficient bit per step; test arbiter for the device jump kernel and
    for spotlight_b200/data/mt19937_jump.npy."""
    ss = np.asarray(key, dtype=np.uint32)
    acc = np.zeros(N, dtype=np.uint32)
    o = 0
    poly = int.from_bytes(np.asarray(poly_words, dtype='<u4').tobytes(), 'little')
    for i in range(poly.bit_length() - 1, -1, -1):
        a0, a1, am = int(acc[o]), int(acc[(o + 1) % N]), int(acc[(o + M) % N])
        y = (a0 & 0x80000000) | (a1 & 0x7FFFFFFF)
        acc[o] = am ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        o = (o + 1) % N
        if (poly >> i) & 1:
            acc ^= np.roll(ss, o)
    return np.roll(acc, -o)
