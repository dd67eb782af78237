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


def test_lazy_adam_state_after_every_step():
    """Step by step: after each minibatch the first moment of every touched row (users and items)
    equals that of a float64 dense Adam driven by the oracle's gradients -- including rows that
    missed steps (their pending steps are replayed BEFORE the forward pass, so the scores of
    step t see the weights dense Adam would have at step t - 1)."""
    from oracle import mf as omf
    from spotlight_b200.factorization.implicit import ImplicitFactorizationModel
    from spotlight_b200.interactions import Interactions
    nsteps, U, I, D, B = 6, 400, 90, 32, 128
    rs = np.random.RandomState(3)
    n = B * nsteps
    users = rs.randint(0, U, n).astype(np.int64)
    items = rs.randint(0, I, n).astype(np.int64)
    inter = Interactions(users.astype(np.int32), items.astype(np.int32), num_users=U, num_items=I)
    m = ImplicitFactorizationModel(loss='bpr', embedding_dim=D, n_iter=1, batch_size=B, use_cuda=True,
                                   random_state=np.random.RandomState(11))
    m._initialize(inter)
    torch.manual_seed(5)
    with torch.no_grad():
        for p in m._net.parameters():
            p.copy_(torch.randn_like(p) * 0.03)
    net = m._net
    P = [p.detach().cpu().numpy().astype(np.float64) for p in
         (net.user_embeddings.weight, net.item_embeddings.weight, net.user_biases.weight, net.item_biases.weight)]
    m._random_state = np.random.RandomState(77)
    negs = np.random.RandomState(77).randint(0, I, n, dtype=np.int64)
    M = [np.zeros_like(p) for p in P]
    V = [np.zeros_like(p) for p in P]
    ud, idv = torch.from_numpy(users).cuda(), torch.from_numpy(items).cuda()
    gaps = set()
    last = {0: np.zeros(U, int), 1: np.zeros(I, int)}
    for t in range(1, nsteps + 1):
        sl = slice((t - 1) * B, t * B)
        m._run_epoch_device(ud[sl], idv[sl])                   # one optimizer step, no flush
        g = omf.mf_step(P[0], P[1], P[2], P[3], users[sl], items[sl], negs[sl], 'bpr', 1, np.float64)
        for k, gr in enumerate((g['dWu'], g['dWi'], g['dbu'], g['dbi'])):
            gr = gr.reshape(P[k].shape)
            M[k] += (gr - M[k]) * 0.1
            V[k] = V[k] * 0.999 + 0.001 * gr * gr
            P[k] -= (1e-2 / (1 - 0.9 ** t)) * (M[k] / (np.sqrt(V[k]) / np.sqrt(1 - 0.999 ** t) + 1e-8))
        for k, prm, ids in ((0, net.user_embeddings.weight, users[sl]),
                            (1, net.item_embeddings.weight, np.concatenate([items[sl], negs[sl]]))):
            rows = np.unique(ids)
            got = m._optimizer.state[prm]['exp_avg'].cpu().numpy()[rows]
            rel = np.abs(got - M[k][rows]).max(1) / np.abs(M[k][rows]).max(1)
            assert rel.max() < 1e-4, (t, k, float(rel.max()))
            assert_close(prm.detach().cpu().numpy()[rows], P[k][rows], 1e-4, what='weights step %d' % t)
            gaps |= set((t - last[k][rows]).tolist())
            last[k][rows] = t
    assert max(gaps) >= 3                                      # rows that missed several steps were exercised


@pytest.mark.parametrize('with_train', [False, True])
def test_mrr_score_blocked_equals_reference_loop(with_train):
    """spotlight_b200.evaluation.mrr_score (user-block GEMM + slb_rank_pairs) against the
    reference's own loop (evaluation.py:38-55: per-user predict, FLOAT_MAX on train items,
    scipy.stats.rankdata), ties included."""
    import scipy.stats as st
    from spotlight_b200 import optim
    from spotlight_b200.evaluation import FLOAT_MAX, mrr_score
    from spotlight_b200.factorization.implicit import ImplicitFactorizationModel
    from spotlight_b200.interactions import Interactions
    rs = np.random.RandomState(1)
    U, I = 300, 120
    train = Interactions(rs.randint(0, U, 4000).astype(np.int32), rs.randint(0, I, 4000).astype(np.int32),
                         num_users=U, num_items=I)
    test = Interactions(rs.randint(0, U, 700).astype(np.int32), rs.randint(0, I, 700).astype(np.int32),
                        num_users=U, num_items=I)
    model = ImplicitFactorizationModel(loss='bpr', embedding_dim=16, n_iter=2, batch_size=256, use_cuda=True,
                                       optimizer_func=optim.fused_adagrad(lr=0.05),
                                       random_state=np.random.RandomState(2))
    model.fit(train)
    with torch.no_grad():                       # exact ties: two identical item rows
        model._net.item_embeddings.weight[7] = model._net.item_embeddings.weight[3]
        model._net.item_biases.weight[7] = model._net.item_biases.weight[3]
    got = mrr_score(model, test, train if with_train else None, user_block=64)
    tcsr, trcsr = test.tocsr(), train.tocsr()
    want = []
    for user_id, row in enumerate(tcsr):
        if not len(row.indices):
            continue
        predictions = -model.predict(user_id)
        if with_train:
            predictions[trcsr[user_id].indices] = FLOAT_MAX
        want.append((1.0 / st.rankdata(predictions)[row.indices]).mean())
    assert got.shape == (len(want),)
    # the block scores come from a GEMM, predict() from the gather kernel: two fp32 summation orders,
    # so a pair of items whose scores differ by ~1e-7 relative may swap ranks for a handful of users
    err = np.abs(got - np.array(want))
    assert np.median(err) < 1e-7 and (err < 1e-6).mean() > 0.9 and err.max() < 2e-2, (np.median(err), err.max())
