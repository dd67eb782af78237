"""Evaluation scoring on the device (reference: spotlight/evaluation.py:9-56).

``mrr_score`` keeps the reference's signature and result -- one score per user with test
interactions, the mean reciprocal *average* rank (``scipy.stats.rankdata`` of the negated
predictions) of the user's test items, known train interactions pushed to the bottom -- but
instead of one ``predict`` + host ranking per user (minutes at 1M users) it scores a block of
users against all items with one GEMM and ranks the (user, test item) pairs with
``slb_rank_pairs``.
"""

import numpy as np
import torch

from spotlight_b200 import _lib, ops

FLOAT_MAX = np.finfo(np.float32).max


def _score_block(model, user_ids):
    """(len(user_ids), num_items) scores of BilinearNet users against every item."""
    net = model._net
    if hasattr(model._optimizer, 'flush'):
        model._optimizer.flush()
    with torch.no_grad():
        u = net.user_embeddings(user_ids)
        out = u @ net.item_embeddings.weight.t()           # plain library GEMM (cuBLAS)
        out += net.user_biases(user_ids).reshape(-1, 1)
        out += net.item_biases.weight.reshape(1, -1)
    return out


def mrr_score(model, test, train=None, user_block=2048):
    """Mean reciprocal rank per user with test interactions (evaluation.py:9-56)."""
    lib = _lib.load()
    test = test.tocsr()
    train = train.tocsr() if train is not None else None
    dev = next(model._net.parameters()).device
    counts = np.diff(test.indptr)
    users = np.nonzero(counts)[0]
    out = np.empty(len(users), dtype=np.float64)
    num_items = model._num_items
    for lo in range(0, len(users), user_block):
        blk = users[lo:lo + user_block]
        scores = _score_block(model, torch.from_numpy(blk.astype(np.int64)).to(dev))
        if train is not None:
            tr = train[blk]
            rows = np.repeat(np.arange(len(blk)), np.diff(tr.indptr))
            if len(rows):
                scores[torch.from_numpy(rows).to(dev), torch.from_numpy(tr.indices.astype(np.int64)).to(dev)] = -float(FLOAT_MAX)
        te = test[blk]
        n_per = np.diff(te.indptr)
        pair_row = torch.from_numpy(np.repeat(np.arange(len(blk)), n_per).astype(np.int64)).to(dev)
        pair_item = torch.from_numpy(te.indices.astype(np.int64)).to(dev)
        ranks = torch.empty(pair_row.numel(), dtype=torch.float32, device=dev)
        _lib.check(lib.slb_rank_pairs(ops._ptr(scores), scores.shape[0], num_items, ops._ptr(pair_row),
                                      ops._ptr(pair_item), pair_row.numel(), ops._ptr(ranks), ops._stream()),
                   'rank_pairs')
        rr = (1.0 / ranks.double()).cpu().numpy()
        ends = np.cumsum(n_per)
        sums = np.add.reduceat(rr, ends - n_per)
        out[lo:lo + len(blk)] = sums / n_per
    return out
