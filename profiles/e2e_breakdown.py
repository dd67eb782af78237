"""Where does model.fit(Interactions) spend its wall time at the bench workload?
Stages timed with a device synchronize on both sides (so they do not overlap here
as they may in fit()); prints one line per stage."""
import argparse, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from spotlight_b200.factorization.implicit import _to_device_narrow
from spotlight_b200.rng import permute_ids, shuffled_order_device
from spotlight_b200.torch_utils import shuffled_order

ap = argparse.ArgumentParser(); ap.add_argument('--batch', type=int, default=524288); ap.add_argument('--steps', type=int, default=100)
x = ap.parse_args()
sys.argv = ['bench.py', '--batch', str(x.batch), '--steps', str(x.steps)]
a = bench.parse()
model = bench.build_model(a, 0)
dev = torch.device('cuda:0'); n = a.batch * a.steps
rs = np.random.RandomState(7)
pu = torch.empty(n, dtype=torch.int32).pin_memory(); pi = torch.empty(n, dtype=torch.int32).pin_memory()
hu, hi = pu.numpy(), pi.numpy(); hu[:] = rs.randint(0, a.users, n); hi[:] = rs.randint(0, a.items, n)
print('from_numpy(view of pinned).is_pinned():', torch.from_numpy(hu).is_pinned())

def timed(name, fn, reps=2):
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        print('%-28s rep %d  %.2f ms' % (name, r, (time.perf_counter() - t0) * 1e3), flush=True)
    return out

ud = timed('ids H2D (users)', lambda: _to_device_narrow(hu, dev))
idv = timed('ids H2D (items)', lambda: _to_device_narrow(hi, dev))
timed('device id range check', lambda: model._check_input(int(ud.max()), int(idv.max())))
timed('host shuffle + H2D', lambda: torch.from_numpy(shuffled_order(n, np.random.RandomState(1))).to(dev).long(), reps=1)
order = timed('device shuffle', lambda: shuffled_order_device(n, np.random.RandomState(1), dev), reps=3)
ref = np.arange(n); np.random.RandomState(1).shuffle(ref)
print('device shuffle == numpy:', bool(np.array_equal(order.cpu().numpy(), ref)))
us = timed('permute_ids', lambda: permute_ids(order, ud, idv))
assert torch.equal(us[0], ud.long()[order]) and torch.equal(us[1], idv.long()[order])
timed('epoch pipeline (K steps)', lambda: model._run_epoch_device(us[0], us[1]))
from spotlight_b200.interactions import Interactions
inter = Interactions(hu, hi, num_users=a.users, num_items=a.items)
timed('model.fit (1 epoch)', lambda: model.fit(inter), reps=3)
# kernel-level view of the device shuffle
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    shuffled_order_device(n, np.random.RandomState(1), dev); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=20, max_name_column_width=48))
