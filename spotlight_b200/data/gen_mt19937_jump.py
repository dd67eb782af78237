"""Generate the MT19937 jump-ahead polynomial table (build-time tool).

The device stream generator (csrc/rng.cu) splits a long run of MT19937 blocks
over many CTAs.  CTA r needs the generator state r*J blocks ahead of the base
state; with T the one-word state transition and phi its minimal polynomial
(degree 19937), T^n S = g(T) S for g = x^n mod phi (Haramoto, Matsumoto,
Nishimura, Panneton, L'Ecuyer: "Efficient jump ahead for F2-linear random
number generators", 2008).  This script computes phi by Berlekamp-Massey on an
output bit sequence and the table g_k = x^(624 * 2^k) mod phi for k = 0..KMAX
by repeated squaring, and writes them as 624 little-endian uint32 words each
(bit i of the polynomial = bit (i % 32) of word i // 32).

A second table serves the one-round generator (slb_mt19937_fill_direct): row r - 1 of
``mt19937_jump_direct.npy`` is x^(624 * J0 * r) mod phi for r = 1..DIRECT_ROWS with
J0 = 2^DIRECT_J0_LOG2 blocks, so CTA r reaches its start state r * J0 blocks ahead
with a single polynomial evaluation instead of log2(r) doubling rounds.

Run:  python spotlight_b200/data/gen_mt19937_jump.py   (under a minute)
"""

import os

import numpy as np

DEG = 19937
KMAX = 26          # jumps up to 2^26 blocks = 4.2e10 words
DIRECT_J0_LOG2 = 8     # fine stride of the direct table: 256 blocks
DIRECT_ROWS = 255      # multiples 1..255 (256 fine slots per coarse slot)

HERE = os.path.dirname(os.path.abspath(__file__))


def mt_words(seed, count):
    """Untempered state-word sequence x[0], x[1], ... of MT19937."""
    x = [0] * (count + 624)
    s = seed & 0xFFFFFFFF
    for i in range(624):
        x[i] = s
        s = (1812433253 * (s ^ (s >> 30)) + i + 1) & 0xFFFFFFFF
    for n in range(count):
        y = (x[n] & 0x80000000) | (x[n + 1] & 0x7FFFFFFF)
        x[n + 624] = x[n + 397] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
    return x


def berlekamp_massey(bits):
    """Connection polynomial C (int, bit i = c_i) and length L of a GF(2) sequence."""
    C, B, L, m = 1, 1, 0, 1
    R = 0                      # bit i of R = s_{n-1-i} (reversed history)
    for n, s in enumerate(bits):
        # discrepancy d = s_n + sum_{i=1..L} c_i s_{n-i}
        d = s ^ (((C >> 1) & R).bit_count() & 1)
        if d:
            T = C
            C ^= B << m
            if 2 * L <= n:
                L, B, m = n + 1 - L, T, 1
            else:
                m += 1
        else:
            m += 1
        R = (R << 1) | s
    return C, L


def reciprocal(C, L):
    out = 0
    for i in range(L + 1):
        if (C >> i) & 1:
            out |= 1 << (L - i)
    return out


_SPREAD = [int(''.join(b + '0' for b in format(v, '08b'))[:-1] or '0', 2) if v else 0 for v in range(256)]


def gf2_square(p):
    """p(x)^2 over GF(2): bit i -> bit 2i."""
    data = p.to_bytes((p.bit_length() + 7) // 8 or 1, 'little')
    out = bytearray(2 * len(data))
    for i, v in enumerate(data):
        w = _SPREAD[v]
        out[2 * i] = w & 0xFF
        out[2 * i + 1] = w >> 8
    return int.from_bytes(out, 'little')


def gf2_mod(p, phi, deg):
    while p.bit_length() > deg:
        p ^= phi << (p.bit_length() - 1 - deg)
    return p


def gf2_mul(a, b):
    """a(x) * b(x) over GF(2)."""
    if a.bit_count() > b.bit_count():
        a, b = b, a
    out = 0
    while a:
        low = a & -a
        out ^= b << (low.bit_length() - 1)
        a ^= low
    return out


def main():
    words = mt_words(5489, 2 * DEG + 1000)
    bits = [w & 1 for w in words[1:2 * DEG + 600]]     # x[0]'s low 31 bits are not state
    C, L = berlekamp_massey(bits)
    assert L == DEG, L
    phi = reciprocal(C, L)
    assert phi.bit_length() == DEG + 1
    table = np.zeros((KMAX + 1, 624), dtype=np.uint32)
    g = gf2_mod(1 << 624, phi, DEG)        # x^624 : one block
    for k in range(KMAX + 1):
        raw = g.to_bytes(624 * 4, 'little')
        table[k] = np.frombuffer(raw, dtype='<u4')
        g = gf2_mod(gf2_square(g), phi, DEG)
    np.save(os.path.join(HERE, 'mt19937_jump.npy'), table)
    base = int.from_bytes(table[DIRECT_J0_LOG2].astype('<u4').tobytes(), 'little')    # x^(624 * J0)
    direct = np.zeros((DIRECT_ROWS, 624), dtype=np.uint32)
    g = base
    for r in range(1, DIRECT_ROWS + 1):
        direct[r - 1] = np.frombuffer(g.to_bytes(624 * 4, 'little'), dtype='<u4')
        g = gf2_mod(gf2_mul(g, base), phi, DEG)
    # the power-of-two multiples must agree with the squaring table
    for k in range(DIRECT_J0_LOG2, DIRECT_J0_LOG2 + 8):
        assert (direct[(1 << (k - DIRECT_J0_LOG2)) - 1] == table[k]).all(), k
    np.save(os.path.join(HERE, 'mt19937_jump_direct.npy'), direct)
    print('phi weight', bin(phi).count('1'), 'table', table.shape)


if __name__ == '__main__':
    main()
