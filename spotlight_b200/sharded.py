"""Multi-GPU training step for the matrix-factorisation hot path.

The reference has no distributed code at all; this is the new design the north
star asks for.  One process per GPU (``torch.distributed``, NCCL over
NVLink/NVSwitch):

* **Partitioning.**  User rows (embedding, bias, optimizer state) are owned by
  contiguous user-id ranges; every interaction of the global minibatch is
  processed by the rank that owns its user, so user gathers and the in-place
  user update are always local.  Item rows are sharded by contiguous index range
  (``owner = id // ceil(num_items / world)``).
* **Forward exchange.**  Each rank buckets the distinct item ids of its local
  batch by owner (``unique_bucket``: one counting pass + scan, ids come out
  ascending = grouped by owner), sends the request lists with an all-to-all,
  owners gather rows + biases from their shard and a second all-to-all returns
  them: each needed row crosses NVLink once per rank per step regardless of how
  often the batch uses it.
* **Local step.**  The fused kernels run on (local user shard, received row
  cache) with the batch's item ids remapped onto the cache; loss and gradients
  are normalised by the *global* batch size so the step equals the single-GPU
  step of the concatenated batch.  User rows are updated in place.
* **Backward exchange.**  The per-distinct-row item gradients (already reduced
  locally, deterministically) travel back with the mirrored all-to-all; each
  owner sums the contributions of its peers in rank order (segmented, no float
  atomics) and applies the row-wise Adagrad update to its shard.
* **Loss.**  One scalar all-reduce.

All collectives are ``all_to_all_single`` / ``all_reduce`` of
``torch.distributed``; the compute pieces are the product's CUDA kernels
(:class:`GpuBackend`).  The routing logic is backend-agnostic so it can be
exercised on CPU with ``gloo`` (tests/test_sharded_cpu.py injects a NumPy
backend there; this module itself never imports the oracle).
"""

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from spotlight_b200 import _lib, ops


class ShardPlan(object):
    """Contiguous range partition of users and items over ``world`` ranks."""

    def __init__(self, num_users, num_items, world):
        self.num_users, self.num_items, self.world = int(num_users), int(num_items), int(world)
        self.uchunk = -(-self.num_users // self.world)
        self.ichunk = -(-self.num_items // self.world)

    def user_range(self, rank):
        lo = min(rank * self.uchunk, self.num_users)
        return lo, min(lo + self.uchunk, self.num_users)

    def item_range(self, rank):
        lo = min(rank * self.ichunk, self.num_items)
        return lo, min(lo + self.ichunk, self.num_items)

    def user_owner(self, user_ids):
        return user_ids // self.uchunk


class GpuBackend(object):
    """The compute pieces of the sharded step on the product's CUDA kernels."""

    def __init__(self, device):
        self.device = torch.device(device)

    def unique_bucket(self, ids, rows, chunk, nparts):
        """distinct ids ascending, inverse map, and per-owner boundaries (host list)."""
        lib = _lib.load()
        ids = ids.contiguous()
        n = ids.numel()
        uniq = torch.empty(min(n, rows), dtype=torch.int64, device=ids.device)
        inverse = torch.empty(n, dtype=torch.int64, device=ids.device)
        counts = torch.empty(nparts + 2, dtype=torch.int64, device=ids.device)
        ws = ops.workspace('uq%d' % rows, lib.slb_unique_workspace_bytes(n, rows), ids.device)
        rc = lib.slb_unique_bucket(ops._ptr(ids), n, rows, chunk, nparts, ops._ptr(uniq),
                                   ops._ptr(inverse), ops._ptr(counts), ops._ptr(ws), ws.numel(),
                                   ops._stream())
        _lib.check(rc, 'unique_bucket')
        host = counts.tolist()                       # the step's one bucketing sync
        return uniq[:host[nparts + 1]], inverse, host[:nparts + 1]

    def unique_bucket_dev(self, ids, rows, chunk, nparts):
        """As unique_bucket, with nothing read back: (uniq [min(n, rows)] of which the first
        counts[nparts + 1] entries are valid, inverse, counts = per-owner boundaries + unique count,
        all on the device)."""
        lib = _lib.load()
        ids = ids.contiguous()
        n = ids.numel()
        uniq = torch.empty(min(n, rows), dtype=torch.int64, device=ids.device)
        inverse = torch.empty(n, dtype=torch.int64, device=ids.device)
        counts = torch.empty(nparts + 2, dtype=torch.int64, device=ids.device)
        ws = ops.workspace('uq%d' % rows, lib.slb_unique_workspace_bytes(n, rows), ids.device)
        rc = lib.slb_unique_bucket(ops._ptr(ids), n, rows, chunk, nparts, ops._ptr(uniq),
                                   ops._ptr(inverse), ops._ptr(counts), ops._ptr(ws), ws.numel(),
                                   ops._stream())
        _lib.check(rc, 'unique_bucket')
        return uniq, inverse, counts

    def gather(self, W, b, local_ids):
        rows = ops.embedding(W, local_ids, [], -1)
        bias = ops.embedding(b.reshape(-1, 1), local_ids, [], -1).reshape(-1)
        return rows, bias

    def local_step(self, st, cache_rows, cache_bias, n_cache, users_local, pos_idx, neg_idx,
                   loss, global_batch, n_neg=1):
        """Fused forward/backward on (user shard, row cache).  Updates the user
        shard in place (row-wise Adagrad); returns (loss share, d cache rows, d cache bias)."""
        lib = _lib.load()
        cap, D = cache_rows.shape
        a = ops.mf_step_args(st.Wu, cache_rows, st.bu, cache_bias, users_local, pos_idx, neg_idx,
                             loss, n_neg)
        loss_out = torch.empty(1, dtype=torch.float32, device=self.device)
        dWi = torch.zeros((cap, D), dtype=torch.float32, device=self.device)
        dbi = torch.zeros(cap, dtype=torch.float32, device=self.device)
        a.loss_out = loss_out.data_ptr()
        a.grad_mode = _lib.GRAD_DENSE
        a.dWi, a.dbi = dWi.data_ptr(), dbi.data_ptr()
        a.opt, a.lr, a.weight_decay, a.eps = _lib.OPT_ADAGRAD, st.lr, 0.0, st.eps
        a.state_Wu, a.state_bu = st.sWu.data_ptr(), st.sbu.data_ptr()
        a.norm_batch, a.opt_users_only = int(global_batch), 1
        need = lib.slb_mf_step_workspace_bytes(a.batch, n_neg, a.loss, a.num_users, a.num_items)
        ws = ops.workspace('mf%d_%d' % (a.num_users, a.num_items), need, self.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        # planned two-kernel step (csrc/mf_v2.cuh): user rows updated in place, the item kernel
        # hands the dense cache-row gradient out for the owners
        need2 = lib.slb_mf_fused_workspace_bytes(a.batch, a.num_users, a.num_items, a.dim) if n_neg == 1 else 0
        if need2:
            fws = ops.workspace('mfv2_%d_%d_%d' % (a.num_users, a.num_items, a.dim), need2,
                                self.device)
            a.fused_workspace, a.fused_workspace_bytes = fws.data_ptr(), fws.numel()
        _lib.check(lib.slb_mf_train_step(ctypes.byref(a), ops._stream()), 'mf_train_step')
        return loss_out.reshape(()), dWi[:n_cache], dbi[:n_cache]

    def owner_update(self, st, local_ids, g_rows, g_bias):
        """Sum the peers' gradient rows per shard row (rank order, deterministic)
        and apply Adagrad to the item shard."""
        rows = st.Wi.shape[0]
        if local_ids.numel() == 0:
            return
        dW = ops.embedding_backward(g_rows.contiguous(), local_ids, [], rows, -1)
        db = ops.embedding_backward(g_bias.reshape(-1, 1).contiguous(), local_ids, [], rows, -1)
        adagrad_dense_(st.Wi, st.sWi, dW, st.lr, st.eps)
        adagrad_dense_(st.bi, st.sbi, db.reshape(-1), st.lr, st.eps)


    # ---- hashed item table (BloomEmbedding, config 4) ----
    def bloom_local_step(self, st, W_full, users_local, items, negs, loss, global_batch):
        """Dense-mode hashed step on (local user shard, full hashed item table, replicated item
        bias): returns (loss share, dWu, dWi, user bias pairs, item bias pairs)."""
        return ops.mf_bloom_step_pairs(st.Wu, W_full, st.bu.reshape(-1, 1), st.bi.reshape(-1, 1), users_local, items,
                                       negs, loss, st.item_seeds, 0, norm_batch=global_batch)

    def adagrad_dense(self, W, S, G, lr, eps):
        _lib.check(_lib.load().slb_adagrad_dense(ops._ptr(W), ops._ptr(S), ops._ptr(G.contiguous()), W.numel(),
                                                 lr, eps, ops._stream()), 'adagrad_dense')

    def bias_sparse_adagrad(self, ids, g, bias, state, lr, eps):
        ops.bias_sparse_apply(ids, g, bias, state, _lib.OPT_ADAGRAD, lr, 0.0, eps)

    # ---- adaptive hinge: scores, loss and score gradients as separate calls ----
    def scores(self, st, cache_rows, cache_bias, u_idx, i_idx):
        return ops.mf_scores(st.Wu, cache_rows, st.bu.reshape(-1, 1), cache_bias.reshape(-1, 1), u_idx, i_idx)

    def adaptive_loss(self, pos, negmat):
        """(mean hinge against the column maxima, d/dpos, d/dneg) -- losses.py:127-166."""
        return ops.pairwise_loss(pos, negmat, None, _lib.LOSS_KIND['adaptive_hinge'])

    def scores_backward(self, st, cache_rows, g, u_idx, i_idx):
        return ops.mf_scores_backward(g, st.Wu, cache_rows, u_idx, i_idx)

    # ---- epoch-level pieces of fit() (all ranks compute the same global stream) ----
    def to_device(self, ids):
        arr = np.ascontiguousarray(ids)
        if arr.dtype not in (np.int32, np.int64):
            arr = arr.astype(np.int64)
        host = torch.from_numpy(arr)
        return host.to(self.device, non_blocking=host.is_pinned())

    def shuffled_order(self, n, random_state):
        from spotlight_b200 import rng
        from spotlight_b200.torch_utils import shuffled_order
        if n >= (1 << 17) and n <= rng.SHUFFLE_DEVICE_MAX:
            return rng.shuffled_order_device(n, random_state, self.device)
        return torch.from_numpy(shuffled_order(n, random_state)).to(self.device).long()

    def permute(self, order, users, items):
        from spotlight_b200 import rng
        return rng.permute_ids(order, users, items)

    def sample(self, num_items, count, random_state):
        from spotlight_b200.sampling import sample_items
        return sample_items(num_items, count, random_state=random_state, device=self.device)

    def upload_sharded(self, ids, rank, world, group=None):
        """Host ids -> the full array on this device with 1/world of the PCIe traffic: every
        rank holds the same host array (single-process semantics), uploads only its slice and
        the slices meet over NVLink (all-gather)."""
        arr = np.ascontiguousarray(ids)
        if arr.dtype not in (np.int32, np.int64):
            arr = arr.astype(np.int64)
        n = arr.shape[0]
        if world == 1 or n < (1 << 16):
            return self.to_device(arr)
        per = -(-n // world)
        lo, hi = min(rank * per, n), min((rank + 1) * per, n)
        host = torch.from_numpy(arr[lo:hi])
        part = torch.zeros(per, dtype=host.dtype, device=self.device)
        part[:hi - lo].copy_(host, non_blocking=host.is_pinned())
        full = torch.empty(per * world, dtype=host.dtype, device=self.device)
        dist.all_gather_into_tensor(full, part, group=group)
        return full[:n]

    def epoch_sampler(self, num_items, random_state, total):
        """Chunked global negative stream on a side stream (device-chained draws, one
        hand-back): draw(count) -> (tensor, event the consumer stream must wait for)."""
        return _GpuEpochSampler(self.device, num_items, random_state, total)

    def seq_local_step(self, E_cache, bias_cache, n_cache, seqs_idx, negs_idx, loss, cnn, norm_count):
        """Fused sequence step on the row cache (ids already remapped onto it; cache
        row 0 is the padding row).  Returns (loss share, dE_cache, dbias_cache, dconv_w, dconv_b)."""
        out = ops.seq_train_step(E_cache, bias_cache.reshape(-1, 1), seqs_idx, negs_idx, loss, 1, cnn,
                                 norm_count=norm_count)
        return (out['loss'], out['dE'][:n_cache], out['dbias'].reshape(-1)[:n_cache],
                out['dconv_w'], out['dconv_b'])


_CHUNK_VALUES = 24 << 20        # negatives per sampler chunk: within the one-round reach of the jump table


class _GpuEpochSampler(object):
    def __init__(self, device, num_items, random_state, total):
        from spotlight_b200 import rng
        from spotlight_b200.factorization.implicit import _side_stream
        self.dev, self.num_items = torch.device(device), int(num_items)
        self.side = _side_stream(self.dev)
        self.out = torch.empty(total, dtype=torch.int64, device=self.dev)
        self.side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(self.side):
            rng.reserve(self.num_items, total, self.dev)     # scratch sized once, before the first draw
            self.stream = rng.DeviceStream(random_state, self.dev)
        self.done = 0

    def draw(self, count):
        lo, self.done = self.done, self.done + count
        with torch.cuda.stream(self.side):
            self.stream.draw(self.num_items, count, out=self.out[lo:self.done])
            ev = torch.cuda.Event()
            ev.record(self.side)
        return self.out[lo:self.done], ev

    def finish(self):
        with torch.cuda.stream(self.side):
            self.stream.finish()
        self.out.record_stream(torch.cuda.current_stream(self.dev))


class _Trace(object):
    """SLB_TRACE=1: host-clock phase times of fit() with a device sync at each mark (diagnostics)."""

    def __init__(self, rank):
        import os
        import time
        self.on = bool(os.environ.get('SLB_TRACE')) and rank == 0
        self.time = time
        self.t = time.perf_counter()

    def __call__(self, what):
        if self.on:
            torch.cuda.synchronize()
            now = self.time.perf_counter()
            print('[trace] %-8s %.2f ms' % (what, (now - self.t) * 1e3), flush=True)
            self.t = now


class _HostEpochSampler(object):
    """Backend-agnostic fallback: one synchronous draw per request (CPU / gloo tests)."""

    def __init__(self, backend, num_items, random_state):
        self.be, self.num_items, self.rs = backend, num_items, random_state

    def draw(self, count):
        return self.be.sample(self.num_items, count, self.rs), None

    def finish(self):
        pass


def adagrad_dense_(W, state, grad, lr, eps):
    """torch.optim.Adagrad update (lr_decay 0) on a small shard; rows with zero
    gradient are unchanged, so this equals the row-wise update of touched rows."""
    state.addcmul_(grad, grad)
    W.addcdiv_(grad, state.sqrt().add_(eps), value=-lr)


class ShardState(object):
    """This rank's parameter shards and Adagrad state."""

    def __init__(self, plan, rank, dim, device, lr=0.05, eps=1e-10, init=None):
        ulo, uhi = plan.user_range(rank)
        ilo, ihi = plan.item_range(rank)
        self.ulo, self.uhi, self.ilo, self.ihi = ulo, uhi, ilo, ihi
        self.lr, self.eps = float(lr), float(eps)
        dev = torch.device(device)
        rows = plan.ichunk              # item shards are padded to the common chunk: the whole-shard
        #                                 exchange (all-gather / reduce-scatter) runs on them directly;
        #                                 rows past ihi - ilo stay zero and are never addressed
        self.Wi = torch.zeros((rows, dim), device=dev)
        self.bi = torch.zeros(rows, device=dev)
        if init is not None:            # slices of full tables (tests / checkpoints)
            Wu, Wi, bu, bi = init
            self.Wu = Wu[ulo:uhi].clone().to(dev)
            self.Wi[:ihi - ilo] = Wi[ilo:ihi].to(dev)
            self.bu = bu[ulo:uhi].reshape(-1).clone().to(dev)
            self.bi[:ihi - ilo] = bi[ilo:ihi].reshape(-1).to(dev)
        else:
            self.Wu = torch.randn((uhi - ulo, dim), device=dev) / dim
            self.Wi[:ihi - ilo] = torch.randn((ihi - ilo, dim), device=dev) / dim
            self.bu = torch.zeros(uhi - ulo, device=dev)
        self.sWu, self.sWi = torch.zeros_like(self.Wu), torch.zeros_like(self.Wi)
        self.sbu, self.sbi = torch.zeros_like(self.bu), torch.zeros_like(self.bi)


class ShardedMF(object):
    """BPR/hinge/pointwise matrix factorisation with range-sharded rows."""

    def __init__(self, plan, state, rank, backend, group=None, cache_capacity=None):
        self.plan, self.st, self.rank, self.backend, self.group = plan, state, rank, backend, group
        self.cache_capacity = cache_capacity
        self.stats = {'rows_requested': 0, 'bytes_a2a': 0}

    def _a2a(self, send, send_counts, recv_counts):
        out = send.new_empty((sum(recv_counts),) + tuple(send.shape[1:]))
        dist.all_to_all_single(out, send.contiguous(), output_split_sizes=list(recv_counts),
                               input_split_sizes=list(send_counts), group=self.group)
        self.stats['bytes_a2a'] += out.numel() * out.element_size()
        return out

    def _dense_exchange_pays(self, local_batch):
        """When a rank's 2*B item draws cover most of the table anyway, the
        per-row routing (bucketing + 3 variable all-to-alls + two host syncs) moves
        as many bytes as shipping whole shards; then all-gather / reduce-scatter of
        the shards is the cheaper exchange."""
        return 2 * local_batch >= self.plan.num_items

    def step(self, users, items, negs, loss, global_batch, exchange='auto', n_neg=1):
        """One training step on this rank's share of the global minibatch.

        ``exchange``: 'a2a' (per-row routing), 'dense' (whole-shard all-gather /
        reduce-scatter) or 'auto'.  ``negs`` holds ``B * n_neg`` ids (adaptive hinge:
        the flat ``randint`` block of the reference, implicit.py:266-275; its user
        pairing ``users[f // n]`` is local because every user of this rank's batch
        is owned by this rank).
        """
        if loss == 'adaptive_hinge' or n_neg != 1:
            # the reference's pairing spans the global minibatch: see step_adaptive
            raise ValueError('adaptive hinge needs the minibatch positions: use step_adaptive')
        if exchange == 'dense' or (exchange == 'auto' and
                                   self._dense_exchange_pays(global_batch // self.plan.world)):
            return self.step_dense(users, items, negs, loss, global_batch, n_neg)
        if exchange == 'a2a_fixed':
            return self.step_a2a_fixed(users, items, negs, loss, global_batch, getattr(self, 'fixed_slots', None))
        return self.step_a2a(users, items, negs, loss, global_batch, n_neg)

    def step_dense(self, users, items, negs, loss, global_batch, n_neg=1):
        """Whole-shard exchange: all-gather the item shards, fused local step on the
        full (transient) item table with raw ids, reduce-scatter the dense item
        gradient back to its owners.  No bucketing, no host synchronisation."""
        plan, st, P = self.plan, self.st, self.plan.world
        chunk, D = plan.ichunk, st.Wi.shape[1]
        dev = users.device
        pad_W, pad_b = st.Wi, st.bi          # shards are stored padded to the common chunk
        full_W = st.Wi.new_empty((P * chunk, D))
        full_b = st.bi.new_empty(P * chunk)
        dist.all_gather_into_tensor(full_W, pad_W, group=self.group)
        dist.all_gather_into_tensor(full_b, pad_b, group=self.group)
        self.stats['bytes_a2a'] += (full_W.numel() + full_b.numel()) * 4
        self.stats['rows_requested'] += P * chunk
        if users.numel():
            loss_share, g_rows, g_bias = self.backend.local_step(
                st, full_W, full_b, P * chunk, users - st.ulo, items, negs, loss, global_batch, n_neg)
        else:                               # none of this minibatch's users live here
            loss_share, g_rows, g_bias = full_b.new_zeros(()), torch.zeros_like(full_W), torch.zeros_like(full_b)
        g_shard = self._reduce_scatter(g_rows.contiguous(), chunk)
        gb_shard = self._reduce_scatter(g_bias.contiguous(), chunk)
        self.stats['bytes_a2a'] += (g_rows.numel() + g_bias.numel()) * 4
        n = st.Wi.shape[0]
        adagrad_dense_(st.Wi, st.sWi, g_shard[:n], st.lr, st.eps)
        adagrad_dense_(st.bi, st.sbi, gb_shard[:n], st.lr, st.eps)
        total = loss_share.detach().clone().reshape(1)
        dist.all_reduce(total, group=self.group)
        return total.reshape(())

    def _reduce_scatter(self, x, chunk):
        out = x.new_empty((chunk,) + tuple(x.shape[1:]))
        try:
            dist.reduce_scatter_tensor(out, x, group=self.group)
        except (RuntimeError, NotImplementedError):          # gloo: sum everywhere, keep our slice
            y = x.clone()
            dist.all_reduce(y, group=self.group)
            out.copy_(y[self.rank * chunk:(self.rank + 1) * chunk])
        return out

    def _fetch_rows(self, ids):
        """Distinct item rows of ``ids`` from their owners (steps 1-3 of the exchange):
        returns (cache_rows, cache_bias, inverse, n_cache, route) with ``ids[k]`` living
        in cache row ``inverse[k]``."""
        plan, st, P = self.plan, self.st, self.plan.world
        dev = ids.device
        if ids.numel():
            uniq, inverse, bounds = self.backend.unique_bucket(ids, plan.num_items, plan.ichunk, P)
        else:                               # nothing of this minibatch lives here: serve peers only
            uniq, inverse, bounds = ids, ids, [0] * (P + 1)
        send_counts = [bounds[p + 1] - bounds[p] for p in range(P)]
        sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        rc = torch.empty(P, dtype=torch.int64, device=dev)
        dist.all_to_all_single(rc, sc, group=self.group)
        recv_counts = rc.tolist()
        req = self._a2a(uniq, send_counts, recv_counts)
        local_req = req - st.ilo
        rows, bias = self.backend.gather(st.Wi, st.bi, local_req)
        n_cache = uniq.numel()
        # fixed capacity (a function of the batch shape only) so the fused step's workspace is reused
        cap = self.cache_capacity or min(ids.numel(), plan.num_items)
        cache_rows = self._a2a(rows, recv_counts, send_counts)
        cache_bias = self._a2a(bias, recv_counts, send_counts)
        if cap > n_cache:                  # fixed-capacity cache keeps the kernel workspace layout stable
            full = cache_rows.new_zeros((cap, cache_rows.shape[1]))
            full[:n_cache] = cache_rows
            fb = cache_bias.new_zeros(cap)
            fb[:n_cache] = cache_bias
            cache_rows, cache_bias = full, fb
        self.stats['rows_requested'] += n_cache
        return cache_rows, cache_bias, inverse, n_cache, (send_counts, recv_counts, local_req)

    def _return_grads(self, route, g_rows, g_bias):
        """Item gradients go home; owners reduce in rank order and update their shard."""
        send_counts, recv_counts, local_req = route
        g_recv = self._a2a(g_rows.contiguous(), send_counts, recv_counts)
        gb_recv = self._a2a(g_bias.contiguous(), send_counts, recv_counts)
        self.backend.owner_update(self.st, local_req, g_recv, gb_recv)

    def _global_loss(self, loss_share):
        total = loss_share.detach().clone().reshape(1)
        dist.all_reduce(total, group=self.group)
        return total.reshape(())

    def step_a2a(self, users, items, negs, loss, global_batch, n_neg=1):
        """Per-row routing (the north-star exchange).

        ``users`` must all be owned by this rank (global ids).  Returns the
        *global* mean loss as a 0-dim tensor (identical on every rank).
        """
        st = self.st
        B = users.numel()
        cache_rows, cache_bias, inverse, n_cache, route = self._fetch_rows(torch.cat([items, negs]))
        # fused local step (user rows updated in place)
        if B:
            loss_share, g_rows, g_bias = self.backend.local_step(
                st, cache_rows, cache_bias, n_cache, users - st.ulo, inverse[:B], inverse[B:], loss,
                global_batch, n_neg)
        else:
            loss_share, g_rows, g_bias = st.bi.new_zeros(()), cache_rows[:0], cache_bias[:0]
        self._return_grads(route, g_rows, g_bias)
        return self._global_loss(loss_share)

    def step_a2a_fixed(self, users, items, negs, loss, global_batch, slots=None):
        """Per-row routing WITHOUT host synchronisation: every rank sends every peer a fixed
        number of request slots (``slots`` per peer; unused slots carry -1), so the all-to-alls
        have equal, host-known splits and the bucket sizes never leave the device.  Traffic is
        padded to the slot capacity; a bucket that does not fit raises ``self.overflow`` (a device
        flag the caller reads once per epoch -- the step's result is then invalid and the epoch
        must be rerun with more slots or the synchronising ``step_a2a``).

        CPU-verified against the float64 oracle (tests/test_sharded_cpu.py, gloo); NOT yet run or
        measured on the B200s -- the round's GPU budget was spent."""
        plan, st, P, be = self.plan, self.st, self.plan.world, self.backend
        dev = users.device
        B = users.numel()
        ids = torch.cat([items, negs])
        n = ids.numel()
        C = int(slots or min(plan.ichunk, (3 * max(n, 1)) // (2 * P) + 1024))
        D = st.Wi.shape[1]
        if not hasattr(self, 'overflow'):
            self.overflow = torch.zeros((), dtype=torch.int64, device=dev)
        if n:
            uniq, inverse, counts = be.unique_bucket_dev(ids, plan.num_items, plan.ichunk, P)
            cap = uniq.numel()
            pos = torch.arange(cap, device=dev)
            valid = pos < counts[P + 1]
            owner = torch.where(valid, torch.div(uniq, plan.ichunk, rounding_mode='floor'), torch.zeros_like(uniq))
            owner = owner.clamp_(0, P - 1)
            rel = pos - counts[:P + 1][owner]
            ok = valid & (rel < C) & (rel >= 0)
            self.overflow += (valid & ~ok).sum()
            slot = torch.where(ok, owner * C + rel, torch.full_like(rel, P * C))      # P*C = a dump slot
            req = torch.full((P * C + 1,), -1, dtype=torch.int64, device=dev)
            req[slot] = torch.where(ok, uniq, torch.full_like(uniq, -1))
            req = req[:P * C].contiguous()
        else:
            cap, inverse = 0, ids
            req = torch.full((P * C,), -1, dtype=torch.int64, device=dev)
        got = torch.empty_like(req)
        dist.all_to_all_single(got, req, group=self.group)
        live = got >= 0
        local = torch.where(live, got - st.ilo, torch.zeros_like(got))
        rows, bias = be.gather(st.Wi, st.bi, local)
        back_rows, back_bias = torch.empty_like(rows), torch.empty_like(bias)
        dist.all_to_all_single(back_rows, rows.contiguous(), group=self.group)
        dist.all_to_all_single(back_bias, bias.contiguous(), group=self.group)
        self.stats['bytes_a2a'] += 2 * (rows.numel() + bias.numel()) * 4 + 8 * req.numel()
        if B:
            take = slot.clamp(max=P * C - 1)
            okf = ok.to(back_rows.dtype)
            cache_rows = back_rows[take] * okf[:, None]
            cache_bias = back_bias[take] * okf
            loss_share, g_rows, g_bias = be.local_step(st, cache_rows, cache_bias, cap, users - st.ulo,
                                                       inverse[:B], inverse[B:], loss, global_batch, 1)
            send_g = g_rows.new_zeros((P * C + 1, D))
            send_gb = g_bias.new_zeros(P * C + 1)
            send_g[slot] = g_rows * okf[:, None]
            send_gb[slot] = g_bias * okf
            send_g, send_gb = send_g[:P * C].contiguous(), send_gb[:P * C].contiguous()
        else:
            loss_share = st.bi.new_zeros(())
            send_g, send_gb = st.Wi.new_zeros((P * C, D)), st.bi.new_zeros(P * C)
        g_recv, gb_recv = torch.empty_like(send_g), torch.empty_like(send_gb)
        dist.all_to_all_single(g_recv, send_g, group=self.group)
        dist.all_to_all_single(gb_recv, send_gb, group=self.group)
        # unused slots carry local row 0 with a zero gradient: they add nothing
        be.owner_update(st, local, g_recv, gb_recv)
        return self._global_loss(loss_share)

    def step_adaptive(self, users, items, negs_block, bpos, batch_users, n_neg):
        """Adaptive hinge on a sharded minibatch, with the reference's pairing.

        The reference scores flat negative f of a minibatch with ``users[f // n]`` and
        reads the result as element ``(k, b) = (f // B, f % B)`` of the ``(n, B)`` matrix
        whose column maxima enter the hinge (implicit.py:266-275, losses.py:127-166).  So
        a negative is *scored* where interaction ``f // n`` lives and *consumed* where
        interaction ``f % B`` lives.  Each rank scores the n-blocks of its own members
        (``negs_block[j*n:(j+1)*n]`` are flats ``bpos[j]*n ..``), the 4-byte scores travel
        to the owners of their columns, the loss and the arg-max gradients are formed
        there, and the gradients travel back the same way.  Rows never move for this:
        only the usual item-row exchange around it.
        """
        plan, st, P = self.plan, self.st, self.plan.world
        be = self.backend
        m, Bg, n = users.numel(), batch_users.numel(), int(n_neg)
        dev = users.device
        cache_rows, cache_bias, inverse, n_cache, route = self._fetch_rows(torch.cat([items, negs_block]))
        ul = users - st.ulo
        ul_rep = ul.repeat_interleave(n)
        # scores of this rank's members and of their n-blocks
        if m:
            pos = be.scores(st, cache_rows, cache_bias, ul, inverse[:m])
            neg = be.scores(st, cache_rows, cache_bias, ul_rep, inverse[m:])
        else:
            pos = st.bi.new_zeros(0)
            neg = st.bi.new_zeros(0)
        # flats -> owners of their columns
        flat = (bpos.repeat_interleave(n) * n + torch.arange(n, device=dev).repeat(m)) if m else bpos
        dest = torch.div(batch_users[flat % Bg], plan.uchunk, rounding_mode='floor') if m else bpos
        order = torch.argsort(dest, stable=True)
        send_counts = torch.bincount(dest, minlength=P)
        recv_counts_t = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts_t, send_counts, group=self.group)
        sc, rc = send_counts.tolist(), recv_counts_t.tolist()
        f_recv = self._a2a(flat[order], sc, rc)
        s_recv = self._a2a(neg[order], sc, rc)
        # the (n, m) matrix of this rank's columns, loss, gradients
        if m:
            lookup = torch.full((Bg,), -1, dtype=torch.int64, device=dev)
            lookup[bpos] = torch.arange(m, device=dev)
            slot = torch.div(f_recv, Bg, rounding_mode='floor') * m + lookup[f_recv % Bg]
            negmat = s_recv.new_empty(n * m)
            negmat[slot] = s_recv
            loss_mean, gp, gn = be.adaptive_loss(pos, negmat.reshape(n, m))
            # hinge gradients are exactly -/+ 1/B_global wherever they are non-zero: emit that
            # value itself (not (1/m) * (m/B), which rounds differently per rank), so that
            # +g and -g meeting on one row cancel exactly as they do in one process
            inv = float(np.float32(1.0) / np.float32(Bg))
            loss_share = loss_mean * (m / float(Bg))
            gp = torch.where(gp != 0, -inv, 0.0).to(torch.float32)
            g_back = torch.where(gn.reshape(-1)[slot] != 0, inv, 0.0).to(torch.float32)
        else:
            loss_share, gp, g_back = st.bi.new_zeros(()), pos, s_recv
        g_sorted = self._a2a(g_back, rc, sc)
        # backward of the scores, with the gradient each score earned at its consumer
        if m:
            g_neg = torch.empty_like(g_sorted)
            g_neg[order] = g_sorted
            dWu, dcache, dbu, dbcache = be.scores_backward(
                st, cache_rows, torch.cat([gp, g_neg]), torch.cat([ul, ul_rep]), inverse)
            adagrad_dense_(st.Wu, st.sWu, dWu, st.lr, st.eps)
            adagrad_dense_(st.bu, st.sbu, dbu.reshape(-1), st.lr, st.eps)
            g_rows, g_bias = dcache[:n_cache], dbcache.reshape(-1)[:n_cache]
        else:
            g_rows, g_bias = cache_rows[:0], cache_bias[:0]
        self._return_grads(route, g_rows, g_bias)
        return self._global_loss(loss_share)


class SeqShardState(object):
    """Item-embedding / item-bias shards (+ replicated conv weights) and Adagrad state."""

    def __init__(self, plan, rank, dim, device, lr=0.05, eps=1e-10, init=None, convs=None):
        ilo, ihi = plan.item_range(rank)
        self.ilo, self.ihi = ilo, ihi
        self.lr, self.eps = float(lr), float(eps)
        dev = torch.device(device)
        if init is not None:
            E, bias = init
            self.Wi = E[ilo:ihi].clone().to(dev)
            self.bi = bias[ilo:ihi].reshape(-1).clone().to(dev)
        else:
            self.Wi = torch.randn((ihi - ilo, dim), device=dev) / dim
            self.bi = torch.zeros(ihi - ilo, device=dev)
            if ilo == 0:
                self.Wi[0] = 0                      # padding row (PADDING_IDX = 0)
        self.sWi, self.sbi = torch.zeros_like(self.Wi), torch.zeros_like(self.bi)
        # conv weights are replicated: list of (weight (D,D,k,1), bias (D,)) tensors
        self.convs = [(w.clone().to(dev), b.clone().to(dev)) for w, b in (convs or [])]
        self.sconvs = [(torch.zeros_like(w), torch.zeros_like(b)) for w, b in self.convs]


class ShardedSeq(object):
    """PoolNet / CNNNet training step with range-sharded item rows.

    Sequences are data-parallel (each rank owns whole sequences); every item row a
    rank's batch touches -- as input, target or negative -- is fetched once per step by
    the same bucket -> all-to-all -> gather -> all-to-all exchange as the MF step, the
    fused sequence kernels run on the row cache, and the per-row gradients return to
    their owners.  The loss is normalised by the *global* number of unmasked positions
    (one scalar all-reduce up front); conv weights are replicated and their gradients
    all-reduced.
    """

    def __init__(self, plan, state, rank, backend, cnn=None, group=None, cache_capacity=None):
        self.plan, self.st, self.rank, self.backend, self.group = plan, state, rank, backend, group
        self.cnn = cnn                      # dict(kernel_width, dilation, nonlinearity, residual) or None
        self.cache_capacity = cache_capacity
        self.stats = {'rows_requested': 0, 'bytes_a2a': 0}

    _a2a = ShardedMF._a2a

    def step(self, seqs, negs, loss):
        plan, st, P = self.plan, self.st, self.plan.world
        B, S = seqs.shape
        dev = seqs.device
        # global mask count first (device scalar; no host sync)
        norm = (seqs != 0).sum().to(torch.int32).reshape(1)
        dist.all_reduce(norm, group=self.group)
        # distinct ids, with the padding id forced in so that it maps to cache row 0
        ids = torch.cat([seqs.reshape(-1), negs.reshape(-1), seqs.new_zeros(1)])
        uniq, inverse, bounds = self.backend.unique_bucket(ids, plan.num_items, plan.ichunk, P)
        send_counts = [bounds[p + 1] - bounds[p] for p in range(P)]
        sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        rc = torch.empty(P, dtype=torch.int64, device=dev)
        dist.all_to_all_single(rc, sc, group=self.group)
        recv_counts = rc.tolist()
        req = self._a2a(uniq, send_counts, recv_counts)
        local_req = req - st.ilo
        rows, bias = self.backend.gather(st.Wi, st.bi, local_req)
        n_cache = uniq.numel()
        # fixed capacity (a function of the batch shape only) so the fused step's workspace is reused
        cap = self.cache_capacity or min(ids.numel(), plan.num_items)
        cache_rows = self._a2a(rows, recv_counts, send_counts)
        cache_bias = self._a2a(bias, recv_counts, send_counts)
        if cap != n_cache:
            full = cache_rows.new_zeros((cap, cache_rows.shape[1]))
            full[:n_cache] = cache_rows
            fb = cache_bias.new_zeros(cap)
            fb[:n_cache] = cache_bias
            cache_rows, cache_bias = full, fb
        self.stats['rows_requested'] += n_cache
        cnn = None
        if self.cnn is not None:
            cnn = dict(self.cnn, weights=[w for w, _ in st.convs], biases=[b for _, b in st.convs])
        n = B * S
        loss_share, g_rows, g_bias, dws, dbs = self.backend.seq_local_step(
            cache_rows, cache_bias, n_cache, inverse[:n].reshape(B, S), inverse[n:2 * n].reshape(B, S),
            loss, cnn, norm)
        g_recv = self._a2a(g_rows.contiguous(), send_counts, recv_counts)
        gb_recv = self._a2a(g_bias.contiguous(), send_counts, recv_counts)
        self.backend.owner_update(st, local_req, g_recv, gb_recv)
        for (w, b), (sw, sb), dw, db in zip(st.convs, st.sconvs, dws, dbs):
            dist.all_reduce(dw, group=self.group)
            dist.all_reduce(db, group=self.group)
            adagrad_dense_(w, sw, dw, st.lr, st.eps)
            adagrad_dense_(b, sb, db, st.lr, st.eps)
        total = loss_share.detach().clone().reshape(1)
        dist.all_reduce(total, group=self.group)
        return total.reshape(())


class ShardedImplicitFactorizationModel(object):
    """``ImplicitFactorizationModel.fit`` on N GPUs (one process per GPU), with the
    *single-process* semantics of the reference loop (factorization/implicit.py:184-252):

    * one global ``RandomState`` stream, advanced identically on every rank: the epoch
      permutation (``shuffle``) and the negatives (``sample_items`` once per minibatch) are
      the reference's, bit for bit;
    * minibatch k is ``shuffled[k*B:(k+1)*B]`` of the *global* data set; each rank trains
      the members whose user it owns and the loss / gradients are those of the whole
      minibatch (:class:`ShardedMF`);
    * ``epoch_loss`` is the mean of the global minibatch losses.

    Every rank is handed the same ``Interactions`` (the global shuffle needs all of it);
    parameters and optimizer state are sharded, never replicated.  Optimizer: row-wise
    Adagrad (``spotlight_b200.optim.fused_adagrad``'s update).  All four losses; adaptive
    hinge keeps the reference's global negative pairing (:meth:`ShardedMF.step_adaptive`).
    """

    def __init__(self, num_users, num_items, rank, world, device, backend=None, loss='bpr',
                 embedding_dim=32, n_iter=10, batch_size=256, learning_rate=0.05, random_state=None,
                 exchange='auto', init=None, group=None, num_negative_samples=5):
        assert loss in ('pointwise', 'bpr', 'hinge', 'adaptive_hinge')
        self._n_neg = int(num_negative_samples) if loss == 'adaptive_hinge' else 1
        self._loss, self._n_iter, self._batch_size = loss, int(n_iter), int(batch_size)
        self._num_users, self._num_items = int(num_users), int(num_items)
        self._random_state = random_state or np.random.RandomState()
        self._exchange = exchange
        self.rank, self.world = rank, world
        self.plan = ShardPlan(num_users, num_items, world)
        self.backend = backend or GpuBackend(device)
        # the reference seeds torch from the model stream at construction (implicit.py:114);
        # the draw is kept so that the stream position matches the single-process model
        seed = int(self._random_state.randint(-10 ** 8, 10 ** 8))
        if init is None:
            torch.manual_seed(seed + 7919 * rank)
        self.state = ShardState(self.plan, rank, embedding_dim, device, lr=learning_rate, init=init)
        self.mf = ShardedMF(self.plan, self.state, rank, self.backend, group=group)
        self.epoch_losses = []

    def fit(self, interactions, verbose=False):
        be = self.backend
        n = len(interactions.user_ids)
        trace = _Trace(self.rank)
        if hasattr(be, 'upload_sharded'):
            users_dev = be.upload_sharded(interactions.user_ids, self.rank, self.world, self.mf.group)
            items_dev = be.upload_sharded(interactions.item_ids, self.rank, self.world, self.mf.group)
        else:
            users_dev = be.to_device(interactions.user_ids)
            items_dev = be.to_device(interactions.item_ids)
        if users_dev.dtype != items_dev.dtype:
            users_dev, items_dev = users_dev.long(), items_dev.long()
        trace('upload')
        if n:
            umax, imax = torch.stack([users_dev.max(), items_dev.max()]).tolist()      # one sync
            if umax >= self._num_users:
                raise ValueError('Maximum user id greater than number of users in model.')
            if imax >= self._num_items:
                raise ValueError('Maximum item id greater than number of items in model.')
        for epoch in range(self._n_iter):
            trace('check')
            order = be.shuffled_order(n, self._random_state)
            trace('shuffle')
            u, i = be.permute(order, users_dev, items_dev)
            del order
            trace('permute')
            epoch_loss = self._run_epoch_device(u, i)
            del u, i
            trace('epoch')
            self.epoch_losses.append(epoch_loss)
            if verbose and self.rank == 0:
                print('Epoch {}: loss {}'.format(epoch, epoch_loss))
            if np.isnan(epoch_loss) or epoch_loss == 0.0:
                raise ValueError('Degenerate epoch loss: {}'.format(epoch_loss))
        return self

    def _run_epoch_device(self, u, i):
        """One epoch over the (already shuffled) global ids ``u`` / ``i``, held identically on
        every rank's device: global negative stream (chunked, on a side stream where the
        backend has one), owner partition, and the sharded steps of this rank's members.
        Returns the epoch loss (mean of the global minibatch losses, implicit.py:240,245)."""
        plan, be, B, nn = self.plan, self.backend, self._batch_size, self._n_neg
        n = u.numel()
        if n == 0:
            return 0.0
        # this rank's members of every minibatch, in minibatch order (no dependence on negatives)
        ulo, uhi = plan.user_range(self.rank)
        mine = torch.nonzero((u >= ulo) & (u < uhi)).reshape(-1)
        edges = torch.arange(0, n + B, B, device=mine.device).clamp_(max=n)
        bounds = torch.searchsorted(mine, edges).tolist()                   # the epoch's one sync
        dense = self._exchange == 'dense' or (self._exchange == 'auto' and
                                              self.mf._dense_exchange_pays(B // plan.world))
        if dense and nn == 1 and isinstance(be, GpuBackend):
            return self._epoch_dense_gpu(u, i, mine, bounds)
        mu, mi = u[mine], i[mine]
        bpos = mine % B
        sampler = (be.epoch_sampler(self._num_items, self._random_state, n * nn)
                   if hasattr(be, 'epoch_sampler') else _HostEpochSampler(be, self._num_items, self._random_state))
        nsteps = len(bounds) - 1
        losses = []
        k, cur = 0, max(1, _CHUNK_VALUES // (B * nn))
        while k < nsteps:
            # one draw per chunk of global minibatches; a minibatch draws len(batch) * n values at
            # once (implicit.py:256-259, 266-275): the n-block of the member at epoch position p
            # is negs[n*p : n*p + n]
            hi_k = min(k + cur, nsteps)
            lo_e, hi_e = k * B, min(hi_k * B, n)
            negs, ev = sampler.draw((hi_e - lo_e) * nn)
            if ev is not None:
                torch.cuda.current_stream(u.device).wait_event(ev)
            negs = negs.reshape(hi_e - lo_e, nn)
            for kk in range(k, hi_k):
                sl = slice(bounds[kk], bounds[kk + 1])
                mn = negs[mine[sl] - lo_e].reshape(-1)
                if self._loss == 'adaptive_hinge':
                    losses.append(self.mf.step_adaptive(mu[sl], mi[sl], mn, bpos[sl], u[kk * B:(kk + 1) * B], nn))
                else:
                    losses.append(self.mf.step(mu[sl], mi[sl], mn, self._loss, min(B, n - kk * B),
                                               self._exchange))
            k = hi_k
        sampler.finish()
        return float(torch.stack(losses).mean()) if losses else 0.0

    # ------------------------------------------------------------------ fast path
    def _buffers(self, key, make):
        cache = self.__dict__.setdefault('_buf_cache', {})
        if key not in cache:
            cache[key] = make()
        return cache[key]

    def _epoch_dense_gpu(self, u, i, mine, bounds):
        """Whole-shard exchange epoch on the product kernels with nothing but launches on the
        host side of the loop: per global minibatch k

          plan stream   gather this rank's members (one kernel), integer plan of the local step
                        (csrc/mf_v2.cuh) -- one step ahead, double buffered
          main stream   all-gather of the item shards -> mf_user_kernel (user rows updated in
                        place) -> mf_item_kernel (dense item gradient) -> reduce-scatter ->
                        Adagrad on the owned shard

        The global negative stream is drawn on the sampler's side stream in chunks; the loss
        shares are all-reduced once per epoch.  Same arithmetic as ShardedMF.step_dense.
        """
        from spotlight_b200.factorization.implicit import _plan_stream
        lib = _lib.load()
        be, st, plan, P = self.backend, self.state, self.plan, self.plan.world
        dev = u.device
        B, n = self._batch_size, u.numel()
        nsteps = len(bounds) - 1
        D, chunk = st.Wi.shape[1], plan.ichunk
        main, pstream = torch.cuda.current_stream(dev), _plan_stream(dev)
        loss_kind = _lib.LOSS_KIND[self._loss]
        maxm = max(1, max(bounds[k + 1] - bounds[k] for k in range(nsteps)))
        cap = 1 << (maxm - 1).bit_length()                    # capacity bucket: buffers are reused across epochs
        buf = self._buffers(('dense', cap), lambda: dict(
            full_W=torch.empty((P * chunk, D), device=dev), full_b=torch.empty(P * chunk, device=dev),
            dW=torch.empty((P * chunk, D), device=dev), db=torch.empty(P * chunk, device=dev),
            gW=torch.empty((chunk, D), device=dev), gb=torch.empty(chunk, device=dev),
            ids=[[torch.empty(cap, dtype=torch.int64, device=dev) for _ in range(3)] for _ in range(2)]))
        U_sh, I_all = st.Wu.shape[0], P * chunk
        fws = ops.workspace('mfv2_%d_%d_%d' % (U_sh, I_all, D), lib.slb_mf_fused_workspace_bytes(cap, U_sh, I_all, D), dev)
        ws = ops.workspace('mf%d_%d' % (U_sh, I_all), lib.slb_mf_step_workspace_bytes(cap, 1, loss_kind, U_sh, I_all), dev)
        assert fws.numel() > 0, 'planned step unavailable for dim %d' % D
        losses = torch.zeros(nsteps, dtype=torch.float32, device=dev)
        sampler = be.epoch_sampler(self._num_items, self._random_state, n)
        # The whole stream is enqueued up front in chunks as large as one jump round reaches
        # (~24 M values): a chunk costs one latency-bound jump round + one fill round whatever its
        # size, and that generator slows down several-fold when it shares SMs with the training
        # kernels -- so as much as possible is drawn before the first step (measured at N = 4: the
        # 1, 2, 4, 8-batch doubling schedule stalled 20 steps for 27 ms in total).
        waits = []                                             # (first step, event) per chunk of negatives
        per = max(1, _CHUNK_VALUES // B)
        k = 0
        while k < nsteps:
            hi_k = min(k + per, nsteps)
            _, ev = sampler.draw(min(hi_k * B, n) - k * B)
            waits.append((k, ev))
            k = hi_k
        import os
        if os.environ.get('SLB_SAMPLER_UPFRONT'):
            # experiment switch: every chunk becomes a dependency of step 0,
            # i.e. the whole epoch's negatives are drawn before the first step and nothing of the
            # generator overlaps the training kernels
            waits = [(0, ev) for _, ev in waits]
        negs_all = sampler.out
        plan_ev = [torch.cuda.Event(), torch.cuda.Event()]
        done_ev = [torch.cuda.Event(), torch.cuda.Event()]
        pstream.wait_stream(main)                              # ids, tables and buffers are ready
        args = [None, None]
        next_wait = [0]

        def make_args(k):
            slot = k & 1
            m = bounds[k + 1] - bounds[k]
            ul, it, ng = buf['ids'][slot]
            a = ops.mf_step_args(st.Wu, buf['full_W'], st.bu, buf['full_b'], ul, it, ng, loss_kind, 1, batch=m)
            a.loss_out = losses[k:k + 1].data_ptr()
            a.grad_mode = _lib.GRAD_DENSE
            a.dWi, a.dbi = buf['dW'].data_ptr(), buf['db'].data_ptr()
            a.opt, a.lr, a.weight_decay, a.eps = _lib.OPT_ADAGRAD, st.lr, 0.0, st.eps
            a.state_Wu, a.state_bu = st.sWu.data_ptr(), st.sbu.data_ptr()
            a.norm_batch, a.opt_users_only = int(min(B, n - k * B)), 1
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
            a.fused_workspace, a.fused_workspace_bytes = fws.data_ptr(), fws.numel()
            return a

        def prep(k):
            slot = k & 1
            m = bounds[k + 1] - bounds[k]
            with torch.cuda.stream(pstream):
                while next_wait[0] < len(waits) and waits[next_wait[0]][0] <= k:
                    pstream.wait_event(waits[next_wait[0]][1])
                    next_wait[0] += 1
                if k >= 2:
                    pstream.wait_event(done_ev[slot])          # the slot's ids / plan are free again
                args[slot] = make_args(k)
                if m:
                    ul, it, ng = buf['ids'][slot]
                    _lib.check(lib.slb_shard_gather_batch(
                        ops._ptr(mine[bounds[k]:]), m, ops._ptr(u), ops._ptr(i), ops._ptr(negs_all), 0, 1,
                        st.ulo, ops._ptr(ul), ops._ptr(it), ops._ptr(ng), ops._stream()), 'shard_gather_batch')
                    _lib.check(lib.slb_mf_train_step_phases(ctypes.byref(args[slot]), 1 | (slot << 8),
                                                            ops._stream()), 'plan')
                plan_ev[slot].record(pstream)

        import os
        import time
        trace = bool(os.environ.get('SLB_TRACE_STEP'))      # diagnostics: device time per phase (events), host time per step
        marks, host_t = [], []

        def mark():
            if trace:
                e = torch.cuda.Event(enable_timing=True)
                e.record(main)
                marks.append(e)

        prep(0)
        for k in range(nsteps):
            t_host = time.perf_counter()
            if k + 1 < nsteps:
                prep(k + 1)
            slot = k & 1
            m = bounds[k + 1] - bounds[k]
            mark()
            dist.all_gather_into_tensor(buf['full_W'], st.Wi, group=self.mf.group)
            dist.all_gather_into_tensor(buf['full_b'], st.bi, group=self.mf.group)
            mark()
            buf['dW'].zero_()
            buf['db'].zero_()
            main.wait_event(plan_ev[slot])
            mark()
            if m:
                _lib.check(lib.slb_mf_train_step_phases(ctypes.byref(args[slot]), 6 | (slot << 8),
                                                        ops._stream()), 'step')
            done_ev[slot].record(main)
            mark()
            dist.reduce_scatter_tensor(buf['gW'], buf['dW'], group=self.mf.group)
            dist.reduce_scatter_tensor(buf['gb'], buf['db'], group=self.mf.group)
            mark()
            _lib.check(lib.slb_adagrad_dense(ops._ptr(st.Wi), ops._ptr(st.sWi), ops._ptr(buf['gW']), chunk * D,
                                             st.lr, st.eps, ops._stream()), 'adagrad')
            _lib.check(lib.slb_adagrad_dense(ops._ptr(st.bi), ops._ptr(st.sbi), ops._ptr(buf['gb']), chunk,
                                             st.lr, st.eps, ops._stream()), 'adagrad')
            mark()
            host_t.append(time.perf_counter() - t_host)
        if trace and self.rank == 0 and nsteps > 4:
            torch.cuda.synchronize()
            names = ['all_gather', 'zero+wait_plan', 'user+item', 'reduce_scatter', 'adagrad', 'gap_to_next']
            acc = [0.0] * 6
            for k in range(2, nsteps - 1):
                ev = marks[6 * k:6 * k + 7]
                for j in range(6):
                    acc[j] += ev[j].elapsed_time(ev[j + 1])
            print('[trace-step] world %d device ms per step:' % self.world,
                  {nm: round(v / (nsteps - 3), 4) for nm, v in zip(names, acc)},
                  'host ms per step %.3f' % (1e3 * sum(host_t[2:]) / len(host_t[2:])), flush=True)
            self.mf.stats['bytes_a2a'] += 2 * (buf['full_W'].numel() + buf['full_b'].numel()) * 4
            self.mf.stats['rows_requested'] += P * chunk
        pstream.wait_stream(main)                              # later plan-stream work follows this epoch
        dist.all_reduce(losses, group=self.mf.group)           # one reduction per epoch: global minibatch losses
        sampler.finish()
        host = losses.cpu().numpy().astype(np.float64)
        if ops.workspace_error_flag(ws):
            raise ValueError('ids out of range reached the device kernels')
        return float(host.mean()) if nsteps else 0.0


class BloomShardState(object):
    """Parameters of BilinearNet(plain users, BloomEmbedding items) on one rank: user rows / user
    bias sharded by user range, the hashed item table (M rows) sharded by row range and padded to
    the common chunk, the item bias (one float per raw item id) REPLICATED -- its forward lookup
    needs 2 values per interaction from arbitrary owners, which would cost a host-synchronised
    all-to-all per step for 4-byte payloads; its replicas are kept identical by applying the same
    all-gathered sparse updates on every rank."""

    def __init__(self, plan, rank, dim, device, num_ids, hashed_rows, num_hash, lr=0.05, eps=1e-10, init=None):
        from spotlight_b200.layers import SEEDS
        dev = torch.device(device)
        self.lr, self.eps = float(lr), float(eps)
        self.ulo, self.uhi = plan.user_range(rank)
        self.M, self.num_ids = int(hashed_rows), int(num_ids)
        self.mchunk = -(-self.M // plan.world)
        self.mlo = min(rank * self.mchunk, self.M)
        self.mhi = min(self.mlo + self.mchunk, self.M)
        self.item_seeds = [int(x) for x in SEEDS[:num_hash]]
        self.Wi = torch.zeros((self.mchunk, dim), device=dev)
        if init is not None:
            Wu, Wi, bu, bi = init
            self.Wu = Wu[self.ulo:self.uhi].clone().to(dev)
            self.bu = bu[self.ulo:self.uhi].reshape(-1).clone().to(dev)
            self.Wi[:self.mhi - self.mlo] = Wi[self.mlo:self.mhi].to(dev)
            self.bi = bi.reshape(-1).clone().to(dev)
        else:
            self.Wu = torch.randn((self.uhi - self.ulo, dim), device=dev) / dim
            self.bu = torch.zeros(self.uhi - self.ulo, device=dev)
            self.Wi[:self.mhi - self.mlo] = torch.randn((self.mhi - self.mlo, dim), device=dev) / dim
            if self.mlo == 0:
                self.Wi[0] = 0                      # padding row of the hashed table
            self.bi = torch.zeros(self.num_ids, device=dev)
        self.sWu, self.sWi = torch.zeros_like(self.Wu), torch.zeros_like(self.Wi)
        self.sbu, self.sbi = torch.zeros_like(self.bu), torch.zeros_like(self.bi)


class ShardedBloomMF(object):
    """Training step of the hashed-item model on N ranks (BASELINE config 4:
    BloomEmbedding 50 M items -> 1 M hashed rows, hinge / bpr / pointwise).

    Interactions are routed to the rank that owns their user (user gathers and updates local).
    The hashed table is range-sharded; a rank's minibatch references 2 * B * H hashed rows -- at
    config 4 sizes a large fraction of all M rows -- so the table travels whole: all-gather of the
    shards, the fused hashed step on the full table (in-register murmur3, layers.py:178-204),
    reduce-scatter of the dense table gradient to the owners, who apply Adagrad.  The id-space
    bias gradients travel as (id, g) pairs: all-gather, then the same sparse update on every
    replica.  Loss: one scalar all-reduce."""

    def __init__(self, plan, state, rank, backend, group=None, pair_capacity=None):
        self.plan, self.st, self.rank, self.backend, self.group = plan, state, rank, backend, group
        self.pair_capacity = pair_capacity
        self.stats = {'bytes_exchanged': 0}

    def step(self, users, items, negs, loss, global_batch):
        st, P, be = self.st, self.plan.world, self.backend
        dev = st.Wi.device
        D = st.Wi.shape[1]
        W_full = st.Wi.new_empty((P * st.mchunk, D))
        dist.all_gather_into_tensor(W_full, st.Wi, group=self.group)
        m = users.numel()
        cap = self.pair_capacity or 2 * int(global_batch)
        ids_pad = torch.zeros(cap, dtype=torch.int64, device=dev)
        g_pad = torch.zeros(cap, dtype=torch.float32, device=dev)
        if m:
            loss_share, dWu, dWi, (iu, gu), (ii, gi) = be.bloom_local_step(st, W_full[:st.M], users - st.ulo, items,
                                                                           negs, loss, global_batch)
            be.adagrad_dense(st.Wu, st.sWu, dWu, st.lr, st.eps)
            be.bias_sparse_adagrad(iu, gu, st.bu, st.sbu, st.lr, st.eps)
            ids_pad[:ii.numel()] = ii
            g_pad[:gi.numel()] = gi
            dW_pad = dWi.new_zeros((P * st.mchunk, D))
            dW_pad[:st.M] = dWi
        else:
            loss_share = st.bu.new_zeros(())
            dW_pad = st.Wi.new_zeros((P * st.mchunk, D))
        g_shard = st.Wi.new_empty((st.mchunk, D))
        try:
            dist.reduce_scatter_tensor(g_shard, dW_pad, group=self.group)
        except (RuntimeError, NotImplementedError):          # gloo: sum everywhere, keep our slice
            y = dW_pad.clone()
            dist.all_reduce(y, group=self.group)
            g_shard.copy_(y[self.rank * st.mchunk:(self.rank + 1) * st.mchunk])
        be.adagrad_dense(st.Wi, st.sWi, g_shard, st.lr, st.eps)
        ids_all = torch.empty(P * cap, dtype=torch.int64, device=dev)
        g_all = torch.empty(P * cap, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(ids_all, ids_pad, group=self.group)
        dist.all_gather_into_tensor(g_all, g_pad, group=self.group)
        be.bias_sparse_adagrad(ids_all, g_all, st.bi, st.sbi, st.lr, st.eps)
        self.stats['bytes_exchanged'] += (W_full.numel() + dW_pad.numel()) * 4 + P * cap * 12
        total = loss_share.detach().clone().reshape(1)
        dist.all_reduce(total, group=self.group)
        return total.reshape(())

