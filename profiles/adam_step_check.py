import numpy as np, torch, sys
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from spotlight_b200.factorization.implicit import ImplicitFactorizationModel
from spotlight_b200.interactions import Interactions
from oracle import mf as omf
nsteps = 5
rs = np.random.RandomState(3)
U, I, D, B = 400, 90, 32, 128
n = B * nsteps
users = rs.randint(0, U, n).astype(np.int64); items = rs.randint(0, I, n).astype(np.int64)
inter = Interactions(users.astype(np.int32), items.astype(np.int32), num_users=U, num_items=I)
dev = torch.device('cuda:0')
for mode in ('one_call', 'step_calls'):
    m = ImplicitFactorizationModel(loss='bpr', embedding_dim=D, n_iter=1, batch_size=B, use_cuda=True, random_state=np.random.RandomState(11))
    m._initialize(inter)
    torch.manual_seed(5)
    with torch.no_grad():
        for p in m._net.parameters(): p.copy_(torch.randn_like(p) * 0.03)
    P = [p.detach().cpu().numpy().astype(np.float64) for p in (m._net.user_embeddings.weight, m._net.item_embeddings.weight, m._net.user_biases.weight, m._net.item_biases.weight)]
    m._random_state = np.random.RandomState(77)
    negs = np.random.RandomState(77).randint(0, I, n, dtype=np.int64)
    M = [np.zeros_like(p) for p in P]; V = [np.zeros_like(p) for p in P]
    ud, idv = torch.from_numpy(users).to(dev), torch.from_numpy(items).to(dev)
    if mode == 'one_call':
        m._run_epoch_device(ud, idv)
    lastu = np.zeros(U, int); lasti = np.zeros(I, int)
    for t in range(1, nsteps + 1):
        sl = slice((t - 1) * B, t * B)
        if mode == 'step_calls':
            m._run_epoch_device(ud[sl], idv[sl])
        g = omf.mf_step(P[0], P[1], P[2], P[3], users[sl], items[sl], negs[sl], 'bpr', 1, np.float64)
        for k, gr in enumerate((g['dWu'], g['dWi'], g['dbu'], g['dbi'])):
            gr = gr.reshape(P[k].shape)
            M[k] += (gr - M[k]) * 0.1; V[k] = V[k] * 0.999 + 0.001 * gr * gr
            P[k] -= (1e-2 / (1 - 0.9 ** t)) * (M[k] / (np.sqrt(V[k]) / np.sqrt(1 - 0.999 ** t) + 1e-8))
        if mode == 'step_calls':
            torch.cuda.synchronize()
            for nm, prm, k, ids, lastv in (('U', m._net.user_embeddings.weight, 0, users[sl], lastu), ('I', m._net.item_embeddings.weight, 1, np.concatenate([items[sl], negs[sl]]), lasti)):
                st = m._optimizer.state[prm]
                rows = np.unique(ids)
                got = st['exp_avg'].cpu().numpy()[rows]; want = M[k][rows]
                rel = np.abs(got - want).max(1) / (np.abs(want).max(1) + 1e-30)
                badr = rows[rel > 1e-3]
                print('step', t, nm, 'touched', len(rows), 'bad', len(badr), 'gaps of bad', sorted(set((t - lastv[badr]).tolist()))[:8], 'gaps of good', sorted(set((t - lastv[rows[rel <= 1e-3]]).tolist()))[:8], 'max rel %.2e' % rel.max())
                lastv[rows] = t
    if mode == 'one_call':
        torch.cuda.synchronize()
        st = m._optimizer.state[m._net.item_embeddings.weight]
        print('one_call: item last', np.unique(st['last'].cpu().numpy()))
