"""MurmurHash3_x86_32 over int32 keys (oracle; test infrastructure only).

``BloomEmbedding`` hashes ids with ``sklearn.utils.murmurhash3_32(x, seed)``
on an int32 array (spotlight/layers.py:183,189), forces the padding id's
hashes to 0 (:184) and reduces with Python's floor-mod by the compressed row
count (:186).  scikit-learn is an unpinned third-party dependency of the
reference (not even in setup.py:13); this restates Appleby's public-domain
MurmurHash3_x86_32 for a 4-byte little-endian key, returning a *signed* int32
exactly like sklearn's default ``positive=False``.  tests/test_oracle_hash.py
pins it against the installed sklearn.
"""

import numpy as np

# spotlight/layers.py:13-20
SEEDS = [
    179424941, 179425457, 179425907, 179426369,
    179424977, 179425517, 179425943, 179426407,
    179424989, 179425529, 179425993, 179426447,
    179425003, 179425537, 179426003, 179426453,
    179425019, 179425559, 179426029, 179426491,
    179425027, 179425579, 179426081, 179426549,
]


def _rotl(x, r):
    return ((x << np.uint32(r)) | (x >> np.uint32(32 - r))).astype(np.uint32)


def murmurhash3_32(keys, seed):
    """Signed int32 hash of each int32 key (4 LE bytes), vectorised."""
    with np.errstate(over='ignore'):
        k = np.asarray(keys).astype(np.int64).astype(np.uint32)
        k = (k * np.uint32(0xCC9E2D51)).astype(np.uint32)
        k = _rotl(k, 15)
        k = (k * np.uint32(0x1B873593)).astype(np.uint32)
        h = np.uint32(int(seed) & 0xFFFFFFFF) ^ k
        h = _rotl(h, 13)
        h = (h * np.uint32(5) + np.uint32(0xE6546B64)).astype(np.uint32)
        h = h ^ np.uint32(4)          # length in bytes
        h ^= h >> np.uint32(16)
        h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
        h ^= h >> np.uint32(13)
        h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
        h ^= h >> np.uint32(16)
    return h.astype(np.uint32).view(np.int32)


def bloom_rows(ids, num_hash_functions, compressed_num_embeddings, padding_idx=0):
    """Hashed row ids, shape ids.shape + (H,), int64.

    layers.py:178-204: ``result[padding_idx] = 0`` is applied to the table
    indexed by id, so id == padding_idx maps to row 0 for every hash; the
    floor-mod keeps rows non-negative for negative int32 hashes.
    """
    ids = np.asarray(ids)
    out = np.empty(ids.shape + (num_hash_functions,), dtype=np.int64)
    for k in range(num_hash_functions):
        h = murmurhash3_32(ids.astype(np.int32), SEEDS[k]).astype(np.int64)
        h = np.where(ids == padding_idx, 0, h)
        out[..., k] = np.mod(h, compressed_num_embeddings)
    return out
