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


@pytest.mark.parametrize('route', ['epoch', 'fused'])
def test_fit_pointwise_default_adam(route, capsys):
    """The reference's default optimizer (dense Adam, implicit.py:143-148).  'epoch': the
    default-constructed model, which now trains with the row-wise lazy-exact Adam
    (csrc/mf_adam.cuh) on the device epoch pipeline; 'fused': stock torch.optim.Adam on the fused
    op's dense gradients.  Both against the reference's recorded trajectory."""
    from spotlight_b200.optim import FusedAdam
    g = load_golden('fit_pointwise_adam')
    func = None if route == 'epoch' else (lambda p: torch.optim.Adam(p, lr=1e-2))
    model, inter = _model(g, 'pointwise', func)
    assert model._route() == route
    assert isinstance(model._optimizer, FusedAdam) == (route == 'epoch')
    losses = _fit_capture(model, inter, capsys)
    assert_close(losses, g['epoch_losses'], 1e-5, what='epoch losses')
    _check_final(model, g, 2e-3)      # Adam normalises by |g|: sign-level sensitivity


@pytest.mark.parametrize('l2', [0.0, 1e-4])
def test_lazy_adam_equals_dense_adam(l2):
    """Row-wise lazy-exact Adam vs torch's dense Adam from the same weights, seed and data, with
    minibatches that touch ~3 % of the rows per step (so a row misses ~30 steps between touches,
    and weight decay -- which moves untouched rows in the dense optimizer -- is replayed too)."""
    from spotlight_b200.factorization.implicit import ImplicitFactorizationModel
    from spotlight_b200.interactions import Interactions
    rs = np.random.RandomState(3)
    U, I, D, n = 4000, 900, 32, 6000
    inter = Interactions(rs.randint(0, U, n).astype(np.int32), rs.randint(0, I, n).astype(np.int32),
                         num_users=U, num_items=I)
    models = []
    for func in (None, lambda p: torch.optim.Adam(p, lr=1e-2, weight_decay=l2)):
        m = ImplicitFactorizationModel(loss='bpr', embedding_dim=D, n_iter=3, batch_size=128, l2=l2,
                                       optimizer_func=func, use_cuda=True,
                                       random_state=np.random.RandomState(11))
        m._initialize(inter)
        models.append(m)
    models[1]._net.load_state_dict(models[0]._net.state_dict())
    with torch.no_grad():
        for m in models:                           # biases start at 0 in the reference: give them values
            m._net.user_biases.weight.copy_(models[0]._net.user_embeddings.weight[:, :1] * 3)
    assert models[0]._route() == 'epoch' and models[1]._route() == 'fused'
    for m in models:
        m.fit(inter)
    assert models[0]._optimizer.steps_taken == 3 * ((n + 127) // 128)
    # 141 Adam steps: an element moves ~lr = 1e-2 per step in the direction of m / sqrt(v), so two
    # fp32 evaluations of the same recurrence (fused multiply-adds here, separate foreach kernels in
    # torch) drift apart where a gradient component is ~0; 5e-4 of the table scale is 3 % of ONE step
    # after a total movement of ~0.5 (measured 1.6e-4).  The formulas are exact: in float64 the two
    # agree to 2e-16, and the per-step state check of profiles/ (exp_avg after every step) is 1e-6.
    for (k, a), (_, b) in zip(models[0]._net.state_dict().items(), models[1]._net.state_dict().items()):
        assert_close(a.cpu().numpy(), b.cpu().numpy(), 5e-4, atol=1e-7, what=k)
    # the optimizer state is dense Adam's too (every row current after fit)
    st0 = models[0]._optimizer.state[models[0]._net.user_embeddings.weight]
    st1 = models[1]._optimizer.state[models[1]._net.user_embeddings.weight]
    assert_close(st0['exp_avg'].cpu().numpy(), st1['exp_avg'].cpu().numpy(), 2e-3, atol=1e-9, what='exp_avg')
    assert_close(st0['exp_avg_sq'].cpu().numpy(), st1['exp_avg_sq'].cpu().numpy(), 2e-3, atol=1e-12, what='exp_avg_sq')
    assert int(st0['last'].min()) == models[0]._optimizer.steps_taken


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
