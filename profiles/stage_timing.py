"""Where does a step go outside the kernels?  Times (CUDA events) the K-step C
pipeline with pre-drawn negatives, the sampler alone, and the full chunked epoch."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from spotlight_b200.sampling import sample_items

ap = argparse.ArgumentParser(); ap.add_argument('--batch', type=int, default=65536); ap.add_argument('--steps', type=int, default=100)
x = ap.parse_args()
sys.argv = ['bench.py', '--batch', str(x.batch), '--steps', str(x.steps)]
a = bench.parse()
model = bench.build_model(a, 0)
dev = torch.device('cuda:0'); B, K = a.batch, a.steps
users = torch.randint(0, a.users, (K * B,), device=dev); items = torch.randint(0, a.items, (K * B,), device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(2):
    e = [ev() for _ in range(4)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e[0].record(); negs = sample_items(a.items, K * B, random_state=model._random_state, device=dev); e[1].record()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    losses = model._fit_epoch_pipeline(users, items, negs, sync=False); e[2].record()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    model._run_epoch_device(users, items); e[3].record(); torch.cuda.synchronize(); t3 = time.perf_counter()
    print('B=%d K=%d  sample: dev %.3f ms wall %.3f ms | pipeline: dev %.3f ms (%.1f us/step) wall %.3f | chunked epoch dev %.3f ms wall %.3f' % (
        B, K, e[0].elapsed_time(e[1]), (t1 - t0) * 1e3, e[1].elapsed_time(e[2]), e[1].elapsed_time(e[2]) / K * 1e3, (t2 - t1) * 1e3,
        e[2].elapsed_time(e[3]), (t3 - t2) * 1e3))
