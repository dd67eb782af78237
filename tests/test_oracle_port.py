"""Pin the torch-CPU port of the reference loop (the timed cpu_baseline)
against the live reference's recorded fit trajectories."""

import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden
from oracle import torch_port


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
