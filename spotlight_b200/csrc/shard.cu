// Per-batch index bucketing for range-sharded embedding rows.
//
// The reference has no multi-GPU path; this is the routing step of the new
// design: the distinct item ids a rank needs this step, in ascending order, are
// automatically grouped by owner (owner = id / chunk), so one counting pass over
// the row space + the segment scan yields the request lists for the all-to-all,
// and `inverse` remaps the batch onto the received row cache.
#include "segindex.cuh"

namespace {

__global__ void __launch_bounds__(256)
uq_count_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t rows, SegIndex seg, int32_t* err) {
    const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += nth) {
        int64_t r = ids[t];
        if (r < 0 || r >= rows) { atomicExch(err, 1); r = 0; }
        atomicAdd(seg.cnt + r, 1);
    }
}

// segment s <-> distinct id seg_row[s]; off[] is reused as the id -> position map
__global__ void __launch_bounds__(256)
uq_emit_kernel(SegIndex seg, int64_t chunk, int nparts, int64_t* __restrict__ uniq, int64_t* counts_out) {
    const int nseg = seg.totals[0];
    const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t s = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; s < nseg; s += nth) {
        const int row = seg.seg_row[s];
        uniq[s] = row;
        seg.off[row] = static_cast<int32_t>(s);
        seg.cnt[row] = 0;                       // restore the zero-at-rest invariant
        // first position owned by each rank: boundary between consecutive distinct ids
        const int64_t own = row / chunk;
        const int64_t prev_own = s == 0 ? -1 : seg.seg_row[s - 1] / chunk;
        for (int64_t p = prev_own + 1; p <= own; ++p) counts_out[p] = s;
        if (s == nseg - 1)
            for (int64_t p = own + 1; p <= nparts; ++p) counts_out[p] = nseg;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counts_out[nparts + 1] = nseg;
        if (nseg == 0) for (int p = 0; p <= nparts; ++p) counts_out[p] = 0;
    }
}

__global__ void __launch_bounds__(256)
uq_inverse_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t rows, SegIndex seg,
                  int64_t* __restrict__ inverse) {
    const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += nth) {
        int64_t r = ids[t];
        if (r < 0 || r >= rows) r = 0;
        inverse[t] = seg.off[r];
    }
}

// users_out[k] = users[pos[k]] - user_lo, items_out[k] = items[pos[k]],
// negs_out[k*n + q] = negs[(pos[k] - neg_base)*n + q]: this rank's members of one global minibatch
__global__ void __launch_bounds__(256)
shard_gather_batch_kernel(const int64_t* __restrict__ pos, int64_t m, const int64_t* __restrict__ users,
                          const int64_t* __restrict__ items, const int64_t* __restrict__ negs, int64_t neg_base,
                          int n_neg, int64_t user_lo, int64_t* __restrict__ users_out,
                          int64_t* __restrict__ items_out, int64_t* __restrict__ negs_out) {
    const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < m; k += nth) {
        const int64_t p = pos[k];
        users_out[k] = users[p] - user_lo;
        items_out[k] = items[p];
        for (int q = 0; q < n_neg; ++q) negs_out[k * n_neg + q] = negs[(p - neg_base) * n_neg + q];
    }
}

// torch.optim.Adagrad (lr_decay 0) on a dense shard; zero gradients leave the element unchanged,
// so this equals the row-wise update of the touched rows
__global__ void __launch_bounds__(256)
adagrad_dense_kernel(float* __restrict__ W, float* __restrict__ S, const float* __restrict__ G, int64_t n,
                     float lr, float eps) {
    const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < n; k += nth) {
        const float g = G[k];
        if (g != 0.f) {
            const float sv = S[k] + g * g;
            S[k] = sv;
            W[k] -= lr * g / (sqrtf(sv) + eps);
        }
    }
}

struct UqLayout { int32_t* flags; SegIndex seg; size_t bytes; };

UqLayout uq_layout(void* base, int64_t n, int64_t rows) {
    WsCarver ws(base);
    UqLayout l;
    l.flags = ws.take<int32_t>(8);
    l.seg = seg_index_carve(ws, rows, n < rows ? n : rows);
    l.bytes = ws.bytes();
    return l;
}

}  // namespace

extern "C" {

size_t slb_unique_workspace_bytes(int64_t n, int64_t rows) { return uq_layout(nullptr, n, rows).bytes; }

int slb_unique_bucket(const int64_t* ids, int64_t n, int64_t rows, int64_t chunk, int32_t nparts,
                      int64_t* uniq, int64_t* inverse, int64_t* counts_out, void* workspace,
                      size_t workspace_bytes, slb_stream_t stream) {
    SLB_REQUIRE(ids && uniq && inverse && counts_out && workspace, "unique_bucket: null pointer");
    SLB_REQUIRE(n > 0 && rows > 0 && chunk > 0 && nparts >= 1, "unique_bucket: bad sizes");
    SLB_REQUIRE(chunk * nparts >= rows, "unique_bucket: chunk * nparts must cover the row space");
    SLB_REQUIRE(rows < (1ll << 31) - SEG_SCAN_TILE && n < (1ll << 31), "unique_bucket: too large");
    UqLayout l = uq_layout(workspace, n, rows);
    if (workspace_bytes < l.bytes) {
        slb_set_error("unique_bucket: workspace too small (%zu < %zu)", workspace_bytes, l.bytes);
        return SLB_ENOSPC;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int sms = slb_sms();
    int g = static_cast<int>((n + 255) / 256);
    if (g > sms * 8) g = sms * 8;
    uq_count_kernel<<<g, 256, 0, st>>>(ids, n, rows, l.seg, l.flags + 4);
    SLB_LAUNCH_CHECK("uq_count_kernel");
    seg_scan_launch(l.seg, l.seg.Rpad, st);
    SLB_LAUNCH_CHECK("seg_scan_kernel");
    uq_emit_kernel<<<g, 256, 0, st>>>(l.seg, chunk, nparts, uniq, counts_out);
    SLB_LAUNCH_CHECK("uq_emit_kernel");
    uq_inverse_kernel<<<g, 256, 0, st>>>(ids, n, rows, l.seg, inverse);
    SLB_LAUNCH_CHECK("uq_inverse_kernel");
    return SLB_OK;
}

int slb_shard_gather_batch(const int64_t* pos, int64_t m, const int64_t* users, const int64_t* items,
                           const int64_t* negs, int64_t neg_base, int32_t n_neg, int64_t user_lo,
                           int64_t* users_out, int64_t* items_out, int64_t* negs_out, slb_stream_t stream) {
    if (m <= 0) return SLB_OK;
    SLB_REQUIRE(pos && users && items && negs && users_out && items_out && negs_out && n_neg >= 1,
                "shard_gather_batch: bad arguments");
    int g = static_cast<int>((m + 255) / 256);
    if (g > slb_sms() * 8) g = slb_sms() * 8;
    shard_gather_batch_kernel<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(pos, m, users, items, negs, neg_base,
                                                                               n_neg, user_lo, users_out, items_out, negs_out);
    SLB_LAUNCH_CHECK("shard_gather_batch_kernel");
    return SLB_OK;
}

int slb_adagrad_dense(float* W, float* state, const float* grad, int64_t n, float lr, float eps,
                      slb_stream_t stream) {
    if (n <= 0) return SLB_OK;
    SLB_REQUIRE(W && state && grad, "adagrad_dense: null pointer");
    int g = static_cast<int>((n + 255) / 256);
    if (g > slb_sms() * 16) g = slb_sms() * 16;
    adagrad_dense_kernel<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(W, state, grad, n, lr, eps);
    SLB_LAUNCH_CHECK("adagrad_dense_kernel");
    return SLB_OK;
}

}  // extern "C"
