// Embedding gathers and their deterministic backward for sm_100a.
//
// Replaces ScaledEmbedding / ZeroEmbedding / BloomEmbedding forward
// (spotlight/layers.py:23-56, 206-244: nn.Embedding lookup; Bloom = index_select
// of the pre-hashed (N, H) table + H-row gather + sum(1)) and
// aten::embedding_dense_backward.  Bloom hashes are computed in registers
// (murmur3 of the id, layers.py:178-204) instead of reading the reference's
// 8*H bytes/id hash table.
#include "segindex.cuh"

namespace {

constexpr int EMB_THREADS = 256;
constexpr int MAX_HASH = 24;  // len(SEEDS), layers.py:13-20

struct HashSpec {
    int32_t H;            // 0 = plain lookup
    int64_t padding_idx;
    uint32_t seeds[MAX_HASH];
};

__device__ __forceinline__ int64_t term_row(const HashSpec& hs, const int64_t* __restrict__ ids,
                                            int64_t t, int64_t rows) {
    if (hs.H == 0) return ids[t];
    return bloom_row(ids[t / hs.H], hs.seeds[t % hs.H], rows, hs.padding_idx);
}

template <int LPR, bool VEC4>
__global__ void __launch_bounds__(EMB_THREADS)
emb_fwd_kernel(const float* __restrict__ W, int64_t rows, int D, const int64_t* __restrict__ ids,
               int64_t n, HashSpec hs, float* __restrict__ out, int32_t* err) {
    constexpr int GROUPS = EMB_THREADS / LPR;
    const int gl = threadIdx.x & (LPR - 1);
    const int fan = hs.H == 0 ? 1 : hs.H;
    for (int64_t b = static_cast<int64_t>(blockIdx.x) * GROUPS + threadIdx.x / LPR; b < n;
         b += static_cast<int64_t>(gridDim.x) * GROUPS) {
        constexpr int STEP = VEC4 ? 4 : 1;
        for (int c = gl * STEP; c < D; c += LPR * STEP) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int k = 0; k < fan; ++k) {
                int64_t r = term_row(hs, ids, b * fan + k, rows);
                if (r < 0 || r >= rows) { if (err) atomicExch(err, 1); r = 0; }
                if (VEC4) {
                    const float4 v = ldg4(W + r * D + c);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                } else {
                    acc.x += __ldg(W + r * D + c);
                }
            }
            if (VEC4) st4(out + b * D + c, acc); else out[b * D + c] = acc.x;
        }
    }
}

__global__ void bloom_rows_kernel(const int64_t* __restrict__ ids, int64_t n, HashSpec hs,
                                  int64_t rows, int64_t* __restrict__ out) {
    const int64_t T = n * hs.H;
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < T;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[t] = term_row(hs, ids, t, rows);
}

__global__ void __launch_bounds__(256)
emb_count_kernel(const int64_t* __restrict__ ids, int64_t T, HashSpec hs, int64_t rows,
                 int32_t* __restrict__ keys, SegIndex seg, int32_t* err) {
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < T;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        int64_t r = term_row(hs, ids, t, rows);
        if (r < 0 || r >= rows) { atomicExch(err, 1); r = 0; }
        keys[t] = static_cast<int32_t>(r);
        atomicAdd(seg.cnt + r, 1);
    }
}

__global__ void __launch_bounds__(256)
emb_fill_kernel(const int32_t* __restrict__ keys, int64_t T, SegIndex seg) {
    seg_rearm(seg);
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < T;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x)
        seg_place(seg, keys[t], static_cast<int32_t>(t));
}

template <int LPR, bool VEC4>
__global__ void __launch_bounds__(EMB_THREADS)
emb_bwd_kernel(const float* __restrict__ dout, int D, int fan, SegIndex seg, int64_t frozen_row,
               float* __restrict__ dW) {
    constexpr int GROUPS = EMB_THREADS / LPR;
    constexpr int CAP = seg_sort_cap(LPR);
    __shared__ int32_t sh_sort[GROUPS * 2 * CAP];
    const int gl = threadIdx.x & (LPR - 1);
    const int gib = threadIdx.x / LPR;
    const unsigned gmask = group_mask(LPR);
    int32_t* sh = sh_sort + gib * 2 * CAP;
    const int nseg = seg.totals[0];
    constexpr int STEP = VEC4 ? 4 : 1;
    for (int64_t s = static_cast<int64_t>(blockIdx.x) * GROUPS + gib; s < nseg;
         s += static_cast<int64_t>(gridDim.x) * GROUPS) {
        const int start = seg.seg_start[s];
        const int len = seg.seg_start[s + 1] - start;
        const int64_t row = seg.seg_row[s];
        if (row == frozen_row) continue;
        for (int c0 = 0; c0 < D; c0 += LPR * STEP) {
            const int c = c0 + gl * STEP;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            seg_visit_sorted<LPR>(seg.members, start, len, gl, gmask, sh, [&](int32_t t) {
                const float* src = dout + static_cast<int64_t>(t / fan) * D;
                if (c < D) {
                    if (VEC4) {
                        const float4 v = ldg4(src + c);
                        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                    } else {
                        acc.x += __ldg(src + c);
                    }
                }
            });
            if (c < D) { if (VEC4) st4(dW + row * D + c, acc); else dW[row * D + c] = acc.x; }
        }
    }
}

int pow2_lanes(int n) {
    int p = 1;
    while (p < n && p < 32) p <<= 1;
    return p;
}

int make_hash(HashSpec& hs, int32_t H, const uint32_t* seeds, int64_t padding_idx) {
    SLB_REQUIRE(H >= 0 && H <= MAX_HASH, "hash_count must be in [0, %d]", MAX_HASH);
    SLB_REQUIRE(H == 0 || seeds != nullptr, "hash seeds missing");
    hs.H = H;
    hs.padding_idx = padding_idx;
    for (int k = 0; k < MAX_HASH; ++k) hs.seeds[k] = k < H ? seeds[k] : 0u;
    return SLB_OK;
}

struct EmbLayout { int32_t* flags; int32_t* keys; SegIndex seg; size_t bytes; };

EmbLayout emb_layout(void* base, int64_t T, int64_t rows) {
    WsCarver ws(base);
    EmbLayout l;
    l.flags = ws.take<int32_t>(8);
    l.seg = seg_index_carve(ws, rows, T);
    l.keys = ws.take<int32_t>(T);
    l.bytes = ws.bytes();
    return l;
}

#define DISPATCH_EMB(lpr, vec4, KERNEL, grid, stream, ...)                                        \
    if (vec4) {                                                                                   \
        switch (lpr) {                                                                            \
            case 1: KERNEL<1, true><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;        \
            case 2: KERNEL<2, true><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;        \
            case 4: KERNEL<4, true><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;        \
            case 8: KERNEL<8, true><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;        \
            case 16: KERNEL<16, true><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;      \
            default: KERNEL<32, true><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;      \
        }                                                                                         \
    } else {                                                                                      \
        switch (lpr) {                                                                            \
            case 1: KERNEL<1, false><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;       \
            case 2: KERNEL<2, false><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;       \
            case 4: KERNEL<4, false><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;       \
            case 8: KERNEL<8, false><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;       \
            case 16: KERNEL<16, false><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;     \
            default: KERNEL<32, false><<<grid, EMB_THREADS, 0, stream>>>(__VA_ARGS__); break;     \
        }                                                                                         \
    }

int grid_for(int64_t work_groups) {
    const int64_t cap = static_cast<int64_t>(slb_sms()) * 8;
    int64_t g = work_groups < cap ? work_groups : cap;
    return g < 1 ? 1 : static_cast<int>(g);
}

}  // namespace

extern "C" {

int slb_embedding_forward(const float* W, int64_t rows, int32_t dim, const int64_t* ids, int64_t n,
                          int32_t hash_count, const uint32_t* seeds, int64_t padding_idx,
                          float* out, slb_stream_t stream) {
    SLB_REQUIRE(W && ids && out, "embedding_forward: null pointer");
    SLB_REQUIRE(rows > 0 && dim > 0, "embedding_forward: bad table shape");
    if (n <= 0) return SLB_OK;
    HashSpec hs;
    const int rc = make_hash(hs, hash_count, seeds, padding_idx);
    if (rc != SLB_OK) return rc;
    const bool vec4 = dim % 4 == 0;
    const int lpr = pow2_lanes(vec4 ? dim / 4 : dim);
    const int grid = grid_for((n + EMB_THREADS / lpr - 1) / (EMB_THREADS / lpr));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DISPATCH_EMB(lpr, vec4, emb_fwd_kernel, grid, st, W, rows, dim, ids, n, hs, out, nullptr);
    SLB_LAUNCH_CHECK("emb_fwd_kernel");
    return SLB_OK;
}

int slb_bloom_rows(const int64_t* ids, int64_t n, int32_t hash_count, const uint32_t* seeds,
                   int64_t rows, int64_t padding_idx, int64_t* rows_out, slb_stream_t stream) {
    SLB_REQUIRE(ids && rows_out && hash_count > 0 && rows > 0, "bloom_rows: bad arguments");
    if (n <= 0) return SLB_OK;
    HashSpec hs;
    const int rc = make_hash(hs, hash_count, seeds, padding_idx);
    if (rc != SLB_OK) return rc;
    bloom_rows_kernel<<<grid_for((n * hash_count + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        ids, n, hs, rows, rows_out);
    SLB_LAUNCH_CHECK("bloom_rows_kernel");
    return SLB_OK;
}

size_t slb_embedding_backward_workspace_bytes(int64_t n_terms, int64_t rows) {
    return emb_layout(nullptr, n_terms, rows).bytes;
}

int slb_embedding_backward(const float* dout, const int64_t* ids, int64_t n, int32_t hash_count,
                           const uint32_t* seeds, int64_t rows, int32_t dim, int64_t frozen_row,
                           float* dW, void* workspace, size_t workspace_bytes, slb_stream_t stream) {
    SLB_REQUIRE(dout && ids && dW && workspace, "embedding_backward: null pointer");
    SLB_REQUIRE(rows > 0 && dim > 0, "embedding_backward: bad table shape");
    if (n <= 0) return SLB_OK;
    HashSpec hs;
    const int rc = make_hash(hs, hash_count, seeds, -1);
    if (rc != SLB_OK) return rc;
    hs.padding_idx = frozen_row;   // Bloom: padding id hashes to row 0 (layers.py:184)
    const int fan = hash_count == 0 ? 1 : hash_count;
    const int64_t T = n * fan;
    SLB_REQUIRE(T < (1ll << 31) && rows < (1ll << 31) - SEG_SCAN_TILE, "embedding_backward: too large");
    EmbLayout l = emb_layout(workspace, T, rows);
    if (workspace_bytes < l.bytes) {
        slb_set_error("embedding_backward: workspace too small (%zu < %zu)", workspace_bytes, l.bytes);
        return SLB_ENOSPC;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int g1 = grid_for((T + 255) / 256);
    emb_count_kernel<<<g1, 256, 0, st>>>(ids, T, hs, rows, l.keys, l.seg, l.flags + 4);
    SLB_LAUNCH_CHECK("emb_count_kernel");
    seg_scan_launch(l.seg, rows, st);
    SLB_LAUNCH_CHECK("seg_scan_kernel");
    emb_fill_kernel<<<g1, 256, 0, st>>>(l.keys, T, l.seg);
    SLB_LAUNCH_CHECK("emb_fill_kernel");
    const bool vec4 = dim % 4 == 0;
    const int lpr = pow2_lanes(vec4 ? dim / 4 : dim);
    const int grid = grid_for((T + EMB_THREADS / lpr - 1) / (EMB_THREADS / lpr));
    // Bloom: the inner table is ScaledEmbedding(M, D, padding_idx=padding_idx)
    // (layers.py:162-164), so the same index is frozen in the compressed table
    DISPATCH_EMB(lpr, vec4, emb_bwd_kernel, grid, st, dout, dim, fan, l.seg, frozen_row, dW);
    SLB_LAUNCH_CHECK("emb_bwd_kernel");
    return SLB_OK;
}

// f3 evaluation scoring: average rank (scipy.stats.rankdata of the NEGATED scores, as
// spotlight/evaluation.py:49 uses it) of selected items within their user's score row:
//   rank = 1 + #(scores > s) + 0.5 * (#(scores == s) - 1).
// One warp per (row, item) pair; the row is L2 resident between the pairs of one user.
static __global__ void __launch_bounds__(256)
rank_pairs_kernel(const float* __restrict__ scores, int64_t n_items, const int64_t* __restrict__ pair_row,
                  const int64_t* __restrict__ pair_item, int64_t n_pairs, float* __restrict__ ranks) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t p = warp; p < n_pairs; p += nwarps) {
        const float* row = scores + pair_row[p] * n_items;
        const float s = row[pair_item[p]];
        int gt = 0, eq = 0;
        for (int64_t k = lane; k < n_items; k += 32) {
            const float v = __ldg(row + k);
            gt += v > s;
            eq += v == s;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            gt += __shfl_xor_sync(0xffffffffu, gt, o);
            eq += __shfl_xor_sync(0xffffffffu, eq, o);
        }
        if (lane == 0) ranks[p] = 1.0f + static_cast<float>(gt) + 0.5f * static_cast<float>(eq - 1);
    }
}

int slb_rank_pairs(const float* scores, int64_t n_rows, int64_t n_items, const int64_t* pair_row,
                   const int64_t* pair_item, int64_t n_pairs, float* ranks, slb_stream_t stream) {
    if (n_pairs <= 0) return SLB_OK;
    SLB_REQUIRE(scores && pair_row && pair_item && ranks && n_rows > 0 && n_items > 0, "rank_pairs: bad arguments");
    const int64_t want = (n_pairs * 32 + 255) / 256;
    const int grid = static_cast<int>(want < static_cast<int64_t>(slb_sms()) * 16 ? want : static_cast<int64_t>(slb_sms()) * 16);
    rank_pairs_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, n_items, pair_row, pair_item, n_pairs, ranks);
    SLB_LAUNCH_CHECK("rank_pairs_kernel");
    return SLB_OK;
}

}  // extern "C"
