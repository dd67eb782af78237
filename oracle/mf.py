"""BilinearNet + implicit losses, closed-form forward/backward (oracle).

TEST INFRASTRUCTURE ONLY.  Restates, in NumPy:

* ``BilinearNet.forward``      spotlight/factorization/representations.py:80-91
* ``pointwise/bpr/hinge/adaptive_hinge`` losses   spotlight/losses.py:40-50,
  82-90, 115-124, 164-166 (masked mean: ``sum(loss*mask)/mask.sum()``)
* the autograd result of ``loss.backward()`` in
  spotlight/factorization/implicit.py:229-243, including the adaptive-hinge
  user/negative misalignment of ``_get_multiple_negative_predictions``
  (implicit.py:266-275: users are repeated ``[u0]*n,[u1]*n,...`` but the flat
  prediction vector is *viewed* as ``(n, B)``).

``dtype`` selects the arithmetic (float32 mimics the reference, float64 gives
a tighter arbiter).  tests/test_oracle_mf.py pins it against golden vectors
produced by the live reference (tests/golden/make_golden.py).
"""

import numpy as np

LOSSES = ('pointwise', 'bpr', 'hinge', 'adaptive_hinge')


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def bilinear_scores(Wu, Wi, bu, bi, users, items, dtype=np.float32):
    """representations.py:80-91: (U[u]*Q[i]).sum(1) + bu[u] + bi[i]."""
    u = Wu[users].astype(dtype)
    q = Wi[items].astype(dtype)
    dot = (u * q).sum(axis=-1, dtype=dtype)
    return dot + bu[users].reshape(dot.shape).astype(dtype) + bi[items].reshape(dot.shape).astype(dtype)


def loss_and_score_grads(kind, pos, neg, mask=None, dtype=np.float32):
    """Loss value and d loss / d pos, d loss / d neg.

    ``neg`` has the shape of ``pos`` except for adaptive_hinge, where it is
    ``(n,) + pos.shape`` and the gradient is routed to the first arg-max over
    axis 0 (torch.max CPU tie-break; losses.py:164).
    """
    pos = pos.astype(dtype)
    neg = neg.astype(dtype)
    if mask is None:
        w = np.full(pos.shape, 1.0 / pos.size, dtype=dtype)
    else:
        m = mask.astype(dtype)
        w = m / m.sum(dtype=dtype)

    if kind == 'adaptive_hinge':
        kstar = np.argmax(neg, axis=0)
        top = np.take_along_axis(neg, kstar[None], axis=0)[0]
        l, gp, gtop = loss_and_score_grads('hinge', pos, top, mask, dtype)
        gn = np.zeros_like(neg)
        np.put_along_axis(gn, kstar[None], gtop[None], axis=0)
        return l, gp, gn

    if kind == 'bpr':
        s = _sigmoid(pos - neg)
        per = 1.0 - s
        gp = -s * (1.0 - s) * w
        gn = -gp
    elif kind == 'hinge':
        z = neg - pos + 1.0
        per = np.maximum(z, 0.0)
        act = (z >= 0.0).astype(dtype)     # clamp backward passes at 0
        gp = -act * w
        gn = act * w
    elif kind == 'pointwise':
        sp = _sigmoid(pos)
        sn = _sigmoid(neg)
        per = (1.0 - sp) + sn
        gp = -sp * (1.0 - sp) * w
        gn = sn * (1.0 - sn) * w
    else:
        raise ValueError(kind)
    if mask is None:
        loss = per.mean(dtype=dtype)
    else:
        loss = (per * m).sum(dtype=dtype) / m.sum(dtype=dtype)
    return dtype(loss), gp.astype(dtype), gn.astype(dtype)


def negative_pairs(users, negs, n_neg, adaptive):
    """(user, item) id pairs the negative predictions are scored with.

    Non-adaptive: pair b = (users[b], negs[b]).  Adaptive (implicit.py:266-275):
    flat index f in [0, B*n) pairs users[f // n] with negs[f]; the (n, B) view
    puts f = k*B + b at [k, b].
    """
    if not adaptive:
        return users, negs
    f = np.arange(users.shape[0] * n_neg)
    return users[f // n_neg], negs.reshape(-1)


def mf_step(Wu, Wi, bu, bi, users, items, negs, loss='bpr', n_neg=1,
            dtype=np.float32):
    """One minibatch: predictions, loss and dense parameter gradients.

    ``negs``: int64 ``[B]`` (or ``[B*n]`` flat for adaptive_hinge, exactly what
    ``sample_items(num_items, B*n)`` returned).
    Returns a dict with pos, neg, loss, dWu, dWi, dbu, dbi (dense, like the
    reference's ``.grad`` with sparse=False).
    """
    adaptive = loss == 'adaptive_hinge'
    B = users.shape[0]
    nu, ni = negative_pairs(users, negs, n_neg, adaptive)
    pos = bilinear_scores(Wu, Wi, bu, bi, users, items, dtype)
    negp = bilinear_scores(Wu, Wi, bu, bi, nu, ni, dtype)
    if adaptive:
        negp = negp.reshape(n_neg, B)
    l, gp, gn = loss_and_score_grads(loss, pos, negp, None, dtype)
    gn_flat = gn.reshape(-1)

    acc = np.float64 if dtype == np.float64 else np.float32
    dWu = np.zeros(Wu.shape, dtype=acc)
    dWi = np.zeros(Wi.shape, dtype=acc)
    dbu = np.zeros(bu.shape, dtype=acc)
    dbi = np.zeros(bi.shape, dtype=acc)
    # index_add in batch order == np.add.at
    np.add.at(dWu, users, gp[:, None] * Wi[items].astype(dtype))
    np.add.at(dWi, items, gp[:, None] * Wu[users].astype(dtype))
    np.add.at(dWu, nu, gn_flat[:, None] * Wi[ni].astype(dtype))
    np.add.at(dWi, ni, gn_flat[:, None] * Wu[nu].astype(dtype))
    np.add.at(dbu.reshape(-1), users, gp)
    np.add.at(dbi.reshape(-1), items, gp)
    np.add.at(dbu.reshape(-1), nu, gn_flat)
    np.add.at(dbi.reshape(-1), ni, gn_flat)
    return dict(pos=pos, neg=negp, loss=l, gp=gp, gn=gn,
                dWu=dWu, dWi=dWi, dbu=dbu, dbi=dbi)


def bloom_embed(W, rows):
    """layers.py:240-241: sum of the H hashed rows.  rows: (..., H)."""
    return W[rows].sum(axis=-2)


def mf_bloom_step(Wu, Wi, bu, bi, users, items, negs, loss, num_hash, pad=0, dtype=np.float64):
    """One minibatch of BilinearNet with a BloomEmbedding ITEM layer (plain user table):
    the item vector is the sum of the H hashed rows (layers.py:206-244), the biases stay
    indexed by the raw ids (representations.py:58-59).  ``Wi`` is the (M, D) hashed table.
    Returns loss and dense gradients dWu, dWi (M rows, padding row 0 frozen), dbu, dbi.
    Non-adaptive losses only (the sharded hashed path, BASELINE config 4, uses hinge)."""
    from oracle.murmur import bloom_rows
    M = Wi.shape[0]
    ri = bloom_rows(np.asarray(items), num_hash, M, pad)
    rj = bloom_rows(np.asarray(negs), num_hash, M, pad)
    Wu_, Wi_ = Wu.astype(dtype), Wi.astype(dtype)
    u = Wu_[users]
    qi, qj = Wi_[ri].sum(1), Wi_[rj].sum(1)
    pos = (u * qi).sum(1) + bu.reshape(-1)[users] + bi.reshape(-1)[items]
    neg = (u * qj).sum(1) + bu.reshape(-1)[users] + bi.reshape(-1)[negs]
    l, gp, gn = loss_and_score_grads(loss, pos.astype(dtype), neg.astype(dtype), None, dtype)
    dWu = np.zeros(Wu.shape, dtype=np.float64)
    dWi = np.zeros(Wi.shape, dtype=np.float64)
    dbu = np.zeros(bu.shape, dtype=np.float64)
    dbi = np.zeros(bi.shape, dtype=np.float64)
    np.add.at(dWu, users, gp[:, None] * qi + gn[:, None] * qj)
    for k in range(num_hash):
        np.add.at(dWi, ri[:, k], gp[:, None] * u)
        np.add.at(dWi, rj[:, k], gn[:, None] * u)
    dWi[0] = 0                                   # the padding row of the compressed table is frozen
    np.add.at(dbu.reshape(-1), users, gp + gn)
    np.add.at(dbi.reshape(-1), items, gp)
    np.add.at(dbi.reshape(-1), negs, gn)
    return dict(loss=l, pos=pos, neg=neg, dWu=dWu, dWi=dWi, dbu=dbu, dbi=dbi)

