"""Pin the torch-CPU port of the reference loop (the timed cpu_baseline)
against the live reference's recorded fit trajectories."""

import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import torch_port

import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('name,loss,opt', [('fit_bpr_sgd', 'bpr', 'sgd'),
                                           ('fit_adaptive_adagrad', 'adaptive_hinge', 'adagrad'),
                                           ('fit_pointwise_adam', 'pointwise', 'adam')])
def test_port_reproduces_reference_fit(name, loss, opt):
    g = load_golden(name)
    torch.set_num_threads(1)
    net = torch_port.PortBilinearNet(int(g['num_users']), int(g['num_items']), int(g['dim']))
    net.load_state_dict({k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('init.')})
    if opt == 'sgd':
        o = torch.optim.SGD(net.parameters(), lr=0.5)
    elif opt == 'adagrad':
        o = torch.optim.Adagrad(net.parameters(), lr=0.05)
    else:
        o = torch.optim.Adam(net.parameters(), weight_decay=0.0, lr=1e-2)
    rs = np.random.RandomState(int(g['seed']))
    rs.randint(-10**8, 10**8)          # the constructor draw, implicit.py:114
    losses = torch_port.fit(net, o, g['users'], g['items'], int(g['num_items']), int(g['batch']),
                            loss, rs, int(g['n_iter']), int(g['n_neg']))
    assert_close(np.array(losses), g['epoch_losses'], 1e-6, what='epoch losses')
    for k, v in net.state_dict().items():
        assert_close(v.numpy(), g['final.' + k], 1e-5, atol=1e-8, what=k)
    st = rs.get_state()
    assert (st[1] == g['rs_key']).all() and st[2] == int(g['rs_pos'])


def test_lazy_adam_scheme_reproduces_reference_default_adam_fit():
    """The row-wise lazy-exact Adam *scheme* the product ships (catch-up of the referenced rows
    before the forward, real step on the touched rows, flush at the end of fit) in NumPy float64
    (oracle/adam.py), driven by the oracle's gradients over the reference's own minibatch sequence,
    against the trajectory the live reference recorded with its default dense Adam
    (spotlight/factorization/implicit.py:143-148; tests/golden/fit_pointwise_adam.npz)."""
    from oracle import mf as omf
    from oracle.adam import LazyAdamTable
    import sharded_common as sc
    g = load_golden('fit_pointwise_adam')
    U, I, B, n_iter = int(g['num_users']), int(g['num_items']), int(g['batch']), int(g['n_iter'])
    epochs, rs = sc.reference_epochs(int(g['seed']), g['users'], g['items'], I, B, n_iter)
    names = ['user_embeddings.weight', 'item_embeddings.weight', 'user_biases.weight', 'item_biases.weight']
    T = [LazyAdamTable(g['init.' + k], lr=1e-2) for k in names]
    t = 0
    losses = []
    for batches in epochs:
        ep = []
        for users, items, negs in batches:
            t += 1
            for tab, rows in ((T[0], users), (T[2], users), (T[1], np.concatenate([items, negs])),
                              (T[3], np.concatenate([items, negs]))):
                tab.catch_up(rows, t - 1)                     # before the forward sees the weights
            r = omf.mf_step(T[0].w, T[1].w, T[2].w, T[3].w, users, items, negs, 'pointwise', 1, np.float64)
            ep.append(float(r['loss']))
            for tab, gr in zip(T, (r['dWu'], r['dWi'], r['dbu'], r['dbi'])):
                gr = gr.reshape(tab.w.shape)
                # the product touches a row when it has a term; rows with an all-zero gradient are
                # simply not touched (same update as a gradient-free step)
                rows = np.nonzero(np.abs(gr).sum(axis=tuple(range(1, gr.ndim))) > 0)[0]
                tab.apply(rows, gr[rows], t)
        losses.append(float(np.mean(ep)))
    for tab in T:
        tab.flush(t)
    assert_close(np.array(losses), g['epoch_losses'], 1e-5, what='epoch losses')
    for tab, k in zip(T, names):
        assert_close(tab.w, g['final.' + k], 2e-3, atol=1e-7, what=k)     # Adam: sign-level sensitivity, as on the GPU
    st = rs.get_state()
    assert (st[1] == g['rs_key']).all() and st[2] == int(g['rs_pos'])
