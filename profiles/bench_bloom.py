"""Secondary measurement: BASELINE.json configs[3] shape on one GPU -- BloomEmbedding
50 M items -> 1 M hashed rows, dim 64, H = 4, hinge loss, plain 1 M-user table, item bias
unhashed (50 M).  Fused hashed-table step, dense gradients (autograd route)."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spotlight_b200 import ops
from spotlight_b200.layers import SEEDS

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=262144); ap.add_argument('--steps', type=int, default=10)
a = ap.parse_args()
dev = torch.device('cuda:0')
U, I, M, D, H, B, K = 1_000_000, 50_000_000, 1_000_000, 64, 4, a.batch, a.steps
torch.manual_seed(0)
Wu = torch.randn(U, D, device=dev) / D
Wi = torch.randn(M, D, device=dev) / D; Wi[0] = 0
bu = torch.zeros(U, 1, device=dev); bi = torch.zeros(I, 1, device=dev)
users = torch.randint(0, U, ((K + 2) * B,), device=dev)
items = torch.randint(0, I, ((K + 2) * B,), device=dev)
negs = torch.randint(0, I, ((K + 2) * B,), device=dev)
def step(k):
    s = slice(k * B, (k + 1) * B)
    return ops.mf_bloom_train_step(Wu, Wi, bu, bi, users[s], items[s], negs[s], 2, 1, [], SEEDS[:H], -1, 0, False)
for k in range(2): step(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(2, 2 + K): out = step(k)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({'config': 'bloom 50M->1M rows, D=64, H=4, hinge, B=%d (dense grads incl. 200 MB item-bias grad)' % B,
                  'ms_per_step': ms, 'interactions_per_s': B / (ms * 1e-3), 'loss': float(out[0])}))
