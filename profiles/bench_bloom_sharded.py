"""Secondary measurement: BASELINE.json configs[3] partitioning on N GPUs (torchrun): BloomEmbedding
50 M items -> 1 M hashed rows (range-sharded, exchanged whole), dim 64, H = 4, hinge, 1 M users
(owner-routed), item bias replicated with all-gathered sparse updates.  Weak scaling: global
minibatch = N x --batch.  Prints one JSON line on rank 0."""
import argparse, json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spotlight_b200.sharded import BloomShardState, GpuBackend, ShardedBloomMF, ShardPlan

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=131072); ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--warmup', type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
U, N, M, D, H = 1_000_000, 50_000_000, 1_000_000, 64, 4
gB, K, W = a.batch * world, a.steps, a.warmup
plan = ShardPlan(U, N, world)
torch.manual_seed(100 + rank)
st = BloomShardState(plan, rank, D, dev, N, M, H, lr=0.05)
model = ShardedBloomMF(plan, st, rank, GpuBackend(dev))
g = torch.Generator(device=dev).manual_seed(77)              # same global ids on every rank
n = (K + W) * gB
users = torch.randint(0, U, (n,), device=dev, generator=g)
items = torch.randint(1, N, (n,), device=dev, generator=g)
negs = torch.randint(0, N, (n,), device=dev, generator=g)
def run(lo, steps):
    last = None
    for k in range(lo, lo + steps):
        s = slice(k * gB, (k + 1) * gB)
        u = users[s]
        mine = torch.nonzero(torch.div(u, plan.uchunk, rounding_mode='floor') == rank).reshape(-1)     # owner routing, in region
        last = model.step(u[mine], items[s][mine], negs[s][mine], 'hinge', gB)
    return last
run(0, W)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); last = run(W, K); e1.record()
dist.barrier(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item()) / K
if rank == 0:
    print(json.dumps({'config': 'bloom 50M->1M hashed rows x%d GPUs (rows range-sharded, whole-table all-gather / reduce-scatter), D=64, H=4, hinge, '
                                'global batch %d' % (world, gB), 'n_gpus': world, 'ms_per_step': ms,
                      'interactions_per_s': gB / (ms * 1e-3), 'loss': float(last),
                      'exchange_gbytes_per_step_per_rank': model.stats['bytes_exchanged'] / (K + W) / 1e9}))
dist.destroy_process_group()
