"""Embedding layers with the reference's names, constructor arguments and
parameter shapes (spotlight/layers.py:13-244), so ``state_dict``s interchange
with the reference.  Lookups and their backward run in the CUDA kernels of
csrc/embed.cu (deterministic segmented scatter instead of
``embedding_dense_backward``); Bloom hashes are computed in registers.
"""

import torch
import torch.nn as nn

from spotlight_b200 import ops

# hash seeds of the reference, spotlight/layers.py:13-20 (24 primes)
SEEDS = [
    179424941, 179425457, 179425907, 179426369,
    179424977, 179425517, 179425943, 179426407,
    179424989, 179425529, 179425993, 179426447,
    179425003, 179425537, 179426003, 179426453,
    179425019, 179425559, 179426029, 179426491,
    179425027, 179425579, 179426081, 179426549
]


class _SparseRowGrad(torch.autograd.Function):
    """Lookup whose weight gradient is an (uncoalesced) sparse COO tensor, the
    contract of ``nn.Embedding(sparse=True)`` that the reference forwards
    (spotlight/factorization/representations.py:49-59)."""

    @staticmethod
    def forward(ctx, W, ids, seeds, padding_idx):
        ctx.save_for_backward(ids)
        ctx.meta = (tuple(W.shape), list(seeds), padding_idx)
        return ops.embedding(W.detach(), ids, seeds, padding_idx)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        shape, seeds, padding_idx = ctx.meta
        flat = ids.reshape(-1)
        if seeds:
            rows = ops.bloom_rows(flat, seeds, shape[0], padding_idx).reshape(-1)
            vals = g.repeat_interleave(len(seeds), dim=0)
        else:
            rows, vals = flat, g
        if padding_idx >= 0:
            vals = vals * (rows != padding_idx).unsqueeze(1).to(vals.dtype)
        return torch.sparse_coo_tensor(rows.unsqueeze(0), vals, shape), None, None, None


def _lookup(weight, ids, seeds, padding_idx, sparse):
    pad = -1 if padding_idx is None else int(padding_idx)
    if sparse and weight.requires_grad and torch.is_grad_enabled():
        return _SparseRowGrad.apply(weight, ids, seeds, pad)
    return ops.embedding(weight, ids, seeds, pad)


class ScaledEmbedding(nn.Embedding):
    """``nn.Embedding`` initialised N(0, 1/embedding_dim), padding row zeroed
    (layers.py:23-37)."""

    def reset_parameters(self):
        self.weight.data.normal_(0, 1.0 / self.embedding_dim)
        if self.padding_idx is not None:
            self.weight.data[self.padding_idx].fill_(0)

    def forward(self, indices):
        out = _lookup(self.weight, indices, [], self.padding_idx, self.sparse)
        return out.view(tuple(indices.shape) + (self.embedding_dim,))


class ZeroEmbedding(nn.Embedding):
    """``nn.Embedding`` initialised to zero; used for biases (layers.py:40-56)."""

    def reset_parameters(self):
        self.weight.data.zero_()
        if self.padding_idx is not None:
            self.weight.data[self.padding_idx].fill_(0)

    def forward(self, indices):
        out = _lookup(self.weight, indices, [], self.padding_idx, self.sparse)
        return out.view(tuple(indices.shape) + (self.embedding_dim,))


class ScaledEmbeddingBag(nn.EmbeddingBag):
    """Kept for API surface (layers.py:59-71); not on the accelerated path."""

    def reset_parameters(self):
        self.weight.data.normal_(0, 1.0 / self.embedding_dim)


class BloomEmbedding(nn.Module):
    """Hashed embedding: every id is represented by the sum of
    ``num_hash_functions`` rows of a ``int(compression_ratio * num_embeddings)``
    row table (layers.py:74-244).

    Row for hash k: ``murmurhash3_32(int32(id), SEEDS[k])`` floor-mod the row
    count; the padding id maps to row 0 for every hash and row ``padding_idx``
    of the table is frozen at zero.  Output shape is ``(batch, seq, dim)`` with
    ``seq = 1`` for 1-d input, as in the reference.

    ``bag=True`` (the reference's EmbeddingBag variant, layers.py:223-236,
    built with offsets that put one element in every bag but the last and
    documented there as performing "very poorly") is not provided.
    """

    def __init__(self, num_embeddings, embedding_dim, compression_ratio=0.2,
                 num_hash_functions=4, bag=False, padding_idx=0):
        super(BloomEmbedding, self).__init__()
        if bag:
            raise NotImplementedError('BloomEmbedding(bag=True) is not supported')
        self.num_embeddings = num_embeddings
        self.embedding_dim = embedding_dim
        self.compression_ratio = compression_ratio
        self.compressed_num_embeddings = int(compression_ratio * num_embeddings)
        self.num_hash_functions = num_hash_functions
        self.padding_idx = padding_idx
        self._bag = bag
        if num_hash_functions > len(SEEDS):
            raise ValueError('Can use at most {} hash functions ({} requested)'
                             .format(len(SEEDS), num_hash_functions))
        self._masks = SEEDS[:self.num_hash_functions]
        self.embeddings = ScaledEmbedding(self.compressed_num_embeddings, self.embedding_dim,
                                          padding_idx=self.padding_idx)

    def __repr__(self):
        return ('<BloomEmbedding (compression_ratio: {}): {}>'
                .format(self.compression_ratio, repr(self.embeddings)))

    def _get_hashed_indices(self, original_indices):
        pad = -1 if self.padding_idx is None else int(self.padding_idx)
        return ops.bloom_rows(original_indices.reshape(-1), self._masks,
                              self.compressed_num_embeddings, pad)

    def forward(self, indices):
        if indices.dim() == 2:
            batch_size, seq_size = indices.size()
        else:
            batch_size, seq_size = indices.size(0), 1
        out = _lookup(self.embeddings.weight, indices.reshape(-1), self._masks,
                      self.padding_idx, self.embeddings.sparse)
        return out.view(batch_size, seq_size, -1)
