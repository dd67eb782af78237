"""2-GPU NCCL test of the sharded step with the product kernels (run with
``gpurun --gpus 2``; skipped on a single GPU)."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, assert_close

sys.path.insert(0, os.path.join(ROOT, 'tests'))
pytestmark = pytest.mark.gpu

SHAPE = (7, 2000, 500, 32, 1024, 3)     # seed, U, I, D, B, steps


def _worker(rank, world, port, loss, q, exchange):
    import sharded_common as sc
    from spotlight_b200.sharded import GpuBackend
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        params, batches = sc.make_problem(*SHAPE)
        dev = torch.device('cuda', rank)
        got, losses, stats = sc.sharded_run(rank, world, params, batches, loss, 0.05, dev,
                                            GpuBackend(dev), cache_capacity=min(2 * SHAPE[4], SHAPE[2]),
                                            exchange=exchange)
        if rank == 0:
            q.put((got, losses, stats))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('loss,exchange', [('bpr', 'a2a'), ('pointwise', 'a2a'), ('bpr', 'dense')])
def test_sharded_gpu_matches_oracle(loss, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import sharded_common as sc
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 3) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, loss, q, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    got, losses, stats = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params, batches = sc.make_problem(*SHAPE)
    ref, ref_losses = sc.oracle_run(params, batches, loss, 0.05)
    assert_close(np.array(losses), np.array(ref_losses), 1e-5, what='losses')
    for a, b, nm in zip(got, ref, ['Wu', 'Wi', 'bu', 'bi']):
        # Adagrad trajectory tolerance: first-touch normalisation amplifies 1e-7 gradient
        # differences on near-cancelling rows (see test_model_gpu / test_sharded_cpu)
        assert_close(a, b, 5e-3, what=nm)


# tanh keeps the scores bounded: with relu at this scale the fp32 sigmoid saturates and a few
# row gradients underflow to exactly 0 where the float64 oracle keeps 1e-16, which Adagrad's
# first-touch normalisation turns into a full lr step (fp32 torch underflows the same way)
_CNN = dict(kernel_width=[3, 3], dilation=[1, 2], nonlinearity='tanh', residual=True)
SEQ_SHAPE = {'pool': (11, 300, 32, 24, 20, 3), 'cnn': (11, 300, 128, 24, 20, 3)}   # seed, I, D, B, S, steps


def _seq_worker(rank, world, port, loss, net, q):
    import sharded_common as sc
    from spotlight_b200.sharded import GpuBackend
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        cnn = _CNN if net == 'cnn' else None
        params, batches = sc.make_seq_problem(*SEQ_SHAPE[net], layers=2 if cnn else 0)
        dev = torch.device('cuda', rank)
        got, losses, stats = sc.seq_sharded_run(rank, world, params, batches, loss, 0.05, dev,
                                                GpuBackend(dev), cnn=cnn)
        if rank == 0:
            q.put((got, losses, stats))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('loss,net', [('bpr', 'pool'), ('pointwise', 'cnn')])
def test_sharded_sequence_gpu_matches_oracle(loss, net):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import sharded_common as sc
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() * 5) % 2000
    procs = [ctx.Process(target=_seq_worker, args=(r, world, port, loss, net, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, losses, stats = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cnn = _CNN if net == 'cnn' else None
    params, batches = sc.make_seq_problem(*SEQ_SHAPE[net], layers=2 if cnn else 0)
    ref, ref_losses = sc.seq_oracle_run(params, batches, loss, 0.05, cnn=cnn)
    assert_close(np.array(losses), np.array(ref_losses), 2e-5, what='losses')
    for k, (a, b) in enumerate(zip(got, ref)):
        assert_close(a, b, 5e-3, what='param%d' % k)      # Adagrad trajectory tolerance, as above


FIT = dict(seed=33, U=3000, I=800, D=32, n=300000, B=16384, n_iter=2)


def _fit_problem():
    import sharded_common as sc
    rs = np.random.RandomState(8)
    params, _ = sc.make_problem(6, FIT['U'], FIT['I'], FIT['D'], 8, 0)
    params = tuple(p * 0.3 for p in params)
    return params, rs.randint(0, FIT['U'], FIT['n']).astype(np.int32), rs.randint(0, FIT['I'], FIT['n']).astype(np.int32)


def _fit_worker(rank, world, port, exchange, q, loss='bpr'):
    import sharded_common as sc
    from spotlight_b200.sharded import GpuBackend
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        params, users, items = _fit_problem()
        dev = torch.device('cuda', rank)
        out = sc.sharded_fit_run(rank, world, params, users, items, loss, dev, GpuBackend(dev),
                                 FIT['seed'], FIT['B'], FIT['n_iter'], exchange, n_neg=4)
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,loss,exchange', [(2, 'bpr', 'a2a'), (2, 'bpr', 'dense'),
                                                 (1, 'adaptive_hinge', 'a2a'), (2, 'adaptive_hinge', 'a2a')])
def test_sharded_fit_equals_single_gpu_fit(world, loss, exchange, capsys):
    """N-GPU fit() vs the single-GPU product fit() from the same RandomState seed and
    weights: same permutation (device shuffle, n >= 2^17), same negatives, same minibatches
    -> same final tables, same final generator state.  The world-1 case runs the whole
    sharded code path (bucketing, all-to-alls with itself, score routing of the adaptive
    hinge) on the product kernels of one GPU."""
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    from spotlight_b200.factorization.implicit import ImplicitFactorizationModel
    from spotlight_b200.interactions import Interactions
    from spotlight_b200.optim import fused_adagrad
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 33500 + (os.getpid() * 7 + world) % 2000
    procs = [ctx.Process(target=_fit_worker, args=(r, world, port, exchange, q, loss)) for r in range(world)]
    for p in procs:
        p.start()
    got, losses, state = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    params, users, items = _fit_problem()
    inter = Interactions(users, items, num_users=FIT['U'], num_items=FIT['I'])
    rs = np.random.RandomState(FIT['seed'])
    one = ImplicitFactorizationModel(loss=loss, embedding_dim=FIT['D'], n_iter=FIT['n_iter'],
                                     batch_size=FIT['B'], use_cuda=True, random_state=rs,
                                     num_negative_samples=4,
                                     optimizer_func=fused_adagrad(lr=0.05))
    one._initialize(inter)
    net = one._net
    with torch.no_grad():
        for prm, val in zip((net.user_embeddings.weight, net.item_embeddings.weight,
                             net.user_biases.weight, net.item_biases.weight), params):
            prm.copy_(torch.from_numpy(val).to(prm.device).reshape(prm.shape))
    one.fit(inter, verbose=True)
    single_losses = [float(line.split('loss')[1]) for line in capsys.readouterr().out.splitlines()
                     if line.startswith('Epoch')]
    ref = [p.detach().cpu().numpy() for p in (net.user_embeddings.weight, net.item_embeddings.weight,
                                              net.user_biases.weight, net.item_biases.weight)]
    assert_close(np.array(losses), np.array(single_losses), 2e-5, what='epoch losses')
    for a, b, nm in zip(got, ref, ['Wu', 'Wi', 'bu', 'bi']):
        if loss == 'adaptive_hinge':
            # This trajectory is chaotic at the row level: in the float64 oracle a 1e-7 relative
            # perturbation of the initial item table moves every user row by more than 1.6e-3
            # (max 0.05 on a 0.32 scale) within these two epochs, while the epoch losses move by
            # < 2e-5 (profiles/adaptive_sensitivity.py).  The two GPU paths score with different
            # kernels (fused tile forward vs mf_scores), i.e. differ by such a perturbation; on
            # the B200 they agree in the epoch losses (2e-5, asserted above) and the generator
            # state, with row differences of the same size as the oracle experiment (max 0.07).
            # The exact semantics are pinned on the CPU against the oracle
            # (tests/test_sharded_cpu.py, worlds 2 and 3).
            # correlation with the oracle under that perturbation: Wu 0.97-0.99, Wi 0.81-0.85,
            # biases 0.35-0.74 (their gradients are mostly exact zeros; what remains is noise)
            floor = {'Wu': 0.8, 'Wi': 0.3}.get(nm)
            assert np.isfinite(a).all()
            if floor is not None:
                assert np.corrcoef(a.reshape(-1), b.reshape(-1))[0, 1] > floor, nm
        else:
            assert_close(a, b.reshape(a.shape), 5e-3, what=nm)   # Adagrad trajectory tolerance, as above
    want = rs.get_state()
    assert np.array_equal(state[1], want[1]) and state[2] == want[2]
    assert len(losses) == FIT['n_iter'] and all(0.0 < v < 1.5 for v in losses)
