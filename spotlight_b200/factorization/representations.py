"""User / item latent representations for factorization models
(reference: spotlight/factorization/representations.py:9-91)."""

import torch.nn as nn

from spotlight_b200 import ops
from spotlight_b200.layers import BloomEmbedding, ScaledEmbedding, ZeroEmbedding


class BilinearNet(nn.Module):
    """Bilinear factorization: score = <user vector, item vector> + user bias
    + item bias.

    Parameter names and shapes match the reference so ``state_dict``s
    interchange: ``user_embeddings.weight (U, D)``, ``item_embeddings.weight
    (I, D)``, ``user_biases.weight (U, 1)``, ``item_biases.weight (I, 1)``.
    Custom ``user_embedding_layer`` / ``item_embedding_layer`` (e.g.
    ``BloomEmbedding``) are accepted as in the reference.
    """

    def __init__(self, num_users, num_items, embedding_dim=32,
                 user_embedding_layer=None, item_embedding_layer=None, sparse=False):
        super(BilinearNet, self).__init__()
        self.embedding_dim = embedding_dim
        self.user_embeddings = (user_embedding_layer if user_embedding_layer is not None
                                else ScaledEmbedding(num_users, embedding_dim, sparse=sparse))
        self.item_embeddings = (item_embedding_layer if item_embedding_layer is not None
                                else ScaledEmbedding(num_items, embedding_dim, sparse=sparse))
        self.user_biases = ZeroEmbedding(num_users, 1, sparse=sparse)
        self.item_biases = ZeroEmbedding(num_items, 1, sparse=sparse)

    def plain_tables(self):
        """True when both sides are un-hashed dense tables without padding, i.e.
        the layout the fused gather-dot and training-step kernels take."""
        def ok(layer):
            return (type(layer) is ScaledEmbedding and layer.padding_idx is None
                    and not layer.sparse and layer.embedding_dim % 4 == 0)
        return (ok(self.user_embeddings) and ok(self.item_embeddings)
                and not self.user_biases.sparse)

    def fused_spec(self):
        """Description of the tables for the fused hashed-table training step, or None
        when a layer is neither a plain ScaledEmbedding nor a BloomEmbedding (or uses
        sparse gradients): ``dict(Wu, Wi, user_seeds, item_seeds, user_pad, item_pad)``."""
        def side(layer):
            if type(layer) is ScaledEmbedding and not layer.sparse and layer.padding_idx is None:
                return layer.weight, [], -1
            if type(layer) is BloomEmbedding and not layer.embeddings.sparse:
                pad = -1 if layer.padding_idx is None else int(layer.padding_idx)
                return layer.embeddings.weight, list(layer._masks), pad
            return None
        u, i = side(self.user_embeddings), side(self.item_embeddings)
        if u is None or i is None or self.user_biases.sparse or self.embedding_dim % 4 != 0:
            return None
        return dict(Wu=u[0], Wi=i[0], user_seeds=u[1], item_seeds=i[1], user_pad=u[2], item_pad=i[2])

    def forward(self, user_ids, item_ids):
        """Predictions for (user, item) pairs, shape ``(batch,)``."""
        if self.plain_tables():
            return ops.mf_scores(self.user_embeddings.weight, self.item_embeddings.weight,
                                 self.user_biases.weight, self.item_biases.weight,
                                 user_ids, item_ids)
        dim = self.embedding_dim
        user_embedding = self.user_embeddings(user_ids).reshape(-1, dim)
        item_embedding = self.item_embeddings(item_ids).reshape(-1, dim)
        user_bias = self.user_biases(user_ids).reshape(-1)
        item_bias = self.item_biases(item_ids).reshape(-1)
        return (user_embedding * item_embedding).sum(1) + user_bias + item_bias
