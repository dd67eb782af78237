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
ap.add_argument('--dense', action='store_true', help='round-1 path: dense gradients, no optimizer')
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
from spotlight_b200 import _lib
states = [torch.zeros_like(x) for x in (Wu, Wi, bu, bi)]
def step(k):
    s = slice(k * B, (k + 1) * B)
    if a.dense:
        return ops.mf_bloom_train_step(Wu, Wi, bu, bi, users[s], items[s], negs[s], 2, 1, [], SEEDS[:H], -1, 0, False)
    return [ops.mf_bloom_train_step_inplace(Wu, Wi, bu, bi, users[s], items[s], negs[s], 'hinge', 1, [], SEEDS[:H],
                                            -1, 0, _lib.OPT_ADAGRAD, 0.05, states=states)]
for k in range(2): step(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(2, 2 + K): out = step(k)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
peak = 6569.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
alg = (1 + 2 * H) * 2 * 4 * D + 40                      # 18R + 40 bytes / interaction at H = 4
print(json.dumps({'config': 'bloom 50M->1M rows, D=64, H=4, hinge, B=%d, %s' % (B, 'dense grads incl. 200 MB item-bias grad' if a.dense else 'fused Adagrad in place (compact rows + sparse bias)'),
                  'ms_per_step': ms, 'interactions_per_s': B / (ms * 1e-3), 'loss': float(out[0]),
                  'algorithmic_bytes_per_interaction': alg, 'roofline_frac': alg * B / (ms * 1e-3) / 1e9 / peak}))
