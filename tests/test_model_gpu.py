"""End-to-end: ImplicitFactorizationModel.fit through every route against
weight trajectories, epoch losses and RandomState positions recorded from the
live reference (tests/golden/fit_*.npz)."""

import numpy as np
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu


def _model(g, loss, optimizer_func, **kw):
    from spotlight_b200.factorization.implicit import ImplicitFactorizationModel
    from spotlight_b200.interactions import Interactions
    inter = Interactions(g['users'], g['items'], num_users=int(g['num_users']),
                         num_items=int(g['num_items']))
    model = ImplicitFactorizationModel(loss=loss, embedding_dim=int(g['dim']),
                                       batch_size=int(g['batch']), n_iter=int(g['n_iter']),
                                       optimizer_func=optimizer_func, use_cuda=True,
                                       num_negative_samples=int(g['n_neg']),
                                       random_state=np.random.RandomState(int(g['seed'])), **kw)
    model._initialize(inter)
    sd = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('init.')}
    model._net.load_state_dict(sd)
    return model, inter


def _check_final(model, g, rtol):
    for k, v in model._net.state_dict().items():
        assert_close(v.cpu().numpy(), g['final.' + k], rtol, atol=1e-7, what=k)
    st = model._random_state.get_state()
    assert (st[1] == g['rs_key']).all() and st[2] == int(g['rs_pos'])


def _fit_capture(model, inter, capsys):
    model.fit(inter, verbose=True)
    lines = [l for l in capsys.readouterr().out.strip().split('\n') if l.startswith('Epoch')]
    return np.array([float(l.split('loss')[1]) for l in lines])


@pytest.mark.parametrize('route', ['fused', 'epoch'])
def test_fit_bpr_sgd(route, capsys):
    from spotlight_b200 import optim
    g = load_golden('fit_bpr_sgd')
    func = (optim.fused_sgd(lr=0.5) if route == 'epoch'
            else (lambda p: torch.optim.SGD(p, lr=0.5)))
    model, inter = _model(g, 'bpr', func)
    assert model._route() == route
    losses = _fit_capture(model, inter, capsys)
    assert_close(losses, g['epoch_losses'], 1e-5, what='epoch losses')
    _check_final(model, g, 1e-5)
    assert_close(model.predict(3), g['predict_user3'], 1e-5, what='predict')


@pytest.mark.parametrize('route', ['fused', 'epoch'])
def test_fit_adaptive_adagrad(route, capsys):
    from spotlight_b200 import optim
    g = load_golden('fit_adaptive_adagrad')
    func = (optim.fused_adagrad(lr=0.05) if route == 'epoch'
            else (lambda p: torch.optim.Adagrad(p, lr=0.05)))
    model, inter = _model(g, 'adaptive_hinge', func)
    assert model._route() == route
    losses = _fit_capture(model, inter, capsys)
    assert_close(losses, g['epoch_losses'], 1e-5, what='epoch losses')
    # adagrad divides by sqrt(sum g^2): tiny grad differences are amplified on
    # first touch, so the trajectory tolerance is looser than the grad tolerance
    _check_final(model, g, 1e-3)


def test_fit_pointwise_default_adam(capsys):
    g = load_golden('fit_pointwise_adam')
    model, inter = _model(g, 'pointwise', None)
    assert model._route() == 'fused'
    losses = _fit_capture(model, inter, capsys)
    assert_close(losses, g['epoch_losses'], 1e-5, what='epoch losses')
    _check_final(model, g, 2e-3)      # Adam normalises by |g|: sign-level sensitivity


def test_fit_resumes_and_pickles(tmp_path):
    from spotlight_b200 import optim
    g = load_golden('fit_bpr_sgd')
    model, inter = _model(g, 'bpr', optim.fused_sgd(lr=0.5))
    model.fit(inter)
    before = model.predict(3)
    path = str(tmp_path / 'm.pt')
    torch.save(model, path)
    again = torch.load(path, weights_only=False)
    assert np.array_equal(again.predict(3), before)
    again.fit(inter)                  # resumes (reference docs/changelog.rst:65)
    assert not np.array_equal(again.predict(3), before)


def test_errors_match_reference():
    from spotlight_b200.factorization.implicit import ImplicitFactorizationModel
    from spotlight_b200.interactions import Interactions
    with pytest.raises(AssertionError):
        ImplicitFactorizationModel(loss='nope')
    inter = Interactions(np.arange(10, dtype=np.int32), np.arange(10, dtype=np.int32))
    with pytest.raises(RuntimeError):
        ImplicitFactorizationModel(use_cuda=False).fit(inter)
    model = ImplicitFactorizationModel(use_cuda=True, n_iter=1)
    model.fit(inter)
    bigger = Interactions(np.arange(20, dtype=np.int32), np.arange(20, dtype=np.int32))
    with pytest.raises(ValueError):
        model.fit(bigger)
