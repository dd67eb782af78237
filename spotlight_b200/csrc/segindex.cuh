// Segment index: the per-batch inverted index that makes the gradient scatter
// deterministic.  Given T "terms" each keyed by a row id in [0, R), build
//   seg_row[s], seg_start[s]  (s = 0..nseg-1, rows ascending), members[]
// so that members[seg_start[s] .. seg_start[s+1]) are the term ids whose key is
// seg_row[s].  One group of lanes then owns each touched row, sums its
// contributions in ascending term order and writes the row once: no float
// atomics, bit-reproducible, replaces aten::embedding_dense_backward's
// index_add (reference: autograd of spotlight/layers.py:23-56 lookups).
//
// Pipeline (all asynchronous on one stream):
//   count   integer atomicAdd into cnt[R]            (fused into the producer)
//   scan    single-pass decoupled look-back scan of cnt -> off, segment list
//   fill    members[off[row] + atomicSub(cnt[row])-1] = term  (restores cnt=0)
//   reduce  consumer kernel; sorts each (tiny) member list, then accumulates
//
// HBM/L2 bytes: cnt read 4R (L2-resident for item tables), everything else
// O(T).  cnt/status/ticket are all-zero at rest.
#pragma once

#include "common.cuh"

constexpr int SEG_SCAN_THREADS = 256;
constexpr int SEG_SCAN_ITEMS = 16;
constexpr int SEG_SCAN_TILE = SEG_SCAN_THREADS * SEG_SCAN_ITEMS;  // 4096
constexpr int SEG_LONG_CTAS = 8;

struct SegIndex {
    int32_t* cnt;        // [Rpad] zero at rest
    int32_t* off;        // [Rpad] exclusive prefix of cnt (valid where cnt > 0)
    int32_t* sid;        // [Rpad] or NULL: segment id of a touched row (planned step, mf_v2.cuh)
    unsigned long long* status;  // [ntiles] look-back words, zero at rest
    int32_t* ticket;     // [1] dynamic tile id, zero at rest
    int32_t* totals;     // [4] nseg, nterms, nsegA, (unused)
    int32_t* seg_row;    // [Tmax]
    int32_t* seg_start;  // [Tmax + 1]
    int32_t* members;    // [Tmax]
    // hot rows (segments longer than long_cap): sorted by a dedicated kernel
    int32_t* long_list;  // [Tmax / 16 + 2] segment ids, totals[3] = count
    int32_t* long_tmp;   // [Tmax] scratch mirror of members
    uint32_t* long_bits; // [SEG_LONG_CTAS][2 * long_words] bitmap + word prefix
    int64_t long_words;  // ceil(Tmax / 32)
    int32_t long_cap;    // segments with more members than this are "long" (0 = feature off)
    int64_t R;           // key space size
    int64_t Rpad;
    int64_t ntiles;
    int64_t Tmax;
};

// Lays the index out in a caller workspace.  Pass base == nullptr to size it.
static inline SegIndex seg_index_carve(WsCarver& ws, int64_t R, int64_t Tmax) {
    SegIndex s;
    s.R = R;
    s.Rpad = (R + SEG_SCAN_TILE - 1) / SEG_SCAN_TILE * SEG_SCAN_TILE;
    s.ntiles = s.Rpad / SEG_SCAN_TILE;
    s.Tmax = Tmax;
    s.cnt = ws.take<int32_t>(s.Rpad);
    s.off = ws.take<int32_t>(s.Rpad);
    s.sid = nullptr;
    s.status = ws.take<unsigned long long>(s.ntiles);
    s.ticket = ws.take<int32_t>(8);
    s.totals = s.ticket + 4;
    s.seg_row = ws.take<int32_t>(Tmax + 1);
    s.seg_start = ws.take<int32_t>(Tmax + 2);
    s.members = ws.take<int32_t>(Tmax + 1);
    s.long_list = ws.take<int32_t>(Tmax / 16 + 2);
    s.long_tmp = ws.take<int32_t>(Tmax + 1);
    s.long_words = (Tmax + 31) / 32;
    s.long_bits = ws.take<uint32_t>(static_cast<size_t>(SEG_LONG_CTAS) * 2 * s.long_words);
    s.long_cap = 0;
    return s;
}

#ifdef __CUDACC__

// status word: [63:62] flag, [61:31] nonzero-count, [30:0] sum
#define SEG_FLAG_AGG 1ull
#define SEG_FLAG_INC 2ull
__device__ __forceinline__ unsigned long long seg_pack(unsigned long long flag, uint32_t sum, uint32_t nz) {
    return (flag << 62) | (static_cast<unsigned long long>(nz) << 31) | sum;
}
__device__ __forceinline__ uint32_t seg_sum(unsigned long long v) { return static_cast<uint32_t>(v & 0x7fffffffull); }
__device__ __forceinline__ uint32_t seg_nz(unsigned long long v) { return static_cast<uint32_t>((v >> 31) & 0x7fffffffull); }
__device__ __forceinline__ unsigned long long seg_flag(unsigned long long v) { return v >> 62; }

// Scan of cnt[0..Rpad) in two launches with no inter-CTA waiting:
//   seg_tilesum_kernel  per-tile (sum, non-zero count) -> status[tile]
//   seg_scan_kernel     each tile sums the aggregates of all earlier tiles (a few
//                       hundred L2-resident words), scans its own 4096 counters and
//                       emits off[], the compact segment list and the totals.
// RA: rows < RA belong to key space A; totals[2] = number of A segments.
static __global__ void __launch_bounds__(SEG_SCAN_THREADS)
seg_tilesum_kernel(SegIndex s) {
    __shared__ unsigned long long sh[SEG_SCAN_THREADS / 32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SEG_SCAN_TILE + threadIdx.x * 4;
    unsigned long long part = 0;      // [63:32] nz, [31:0] sum
#pragma unroll
    for (int i = 0; i < SEG_SCAN_ITEMS / 4; ++i) {
        const int4 v = *reinterpret_cast<const int4*>(s.cnt + base + i * SEG_SCAN_THREADS * 4);
        part += static_cast<unsigned long long>(static_cast<uint32_t>(v.x + v.y + v.z + v.w)) |
                (static_cast<unsigned long long>((v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0)) << 32);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
#pragma unroll
        for (int w = 0; w < SEG_SCAN_THREADS / 32; ++w) tot += sh[w];
        s.status[blockIdx.x] = tot;
    }
}

static __global__ void __launch_bounds__(SEG_SCAN_THREADS)
seg_scan_kernel(SegIndex s, int64_t RA) {
    __shared__ uint32_t sh_wsum[SEG_SCAN_THREADS / 32], sh_wnz[SEG_SCAN_THREADS / 32];
    __shared__ unsigned long long sh_part[SEG_SCAN_THREADS / 32];
    __shared__ uint32_t sh_prefix_sum, sh_prefix_nz;
    const int tile = blockIdx.x;
    const int64_t base = static_cast<int64_t>(tile) * SEG_SCAN_TILE + threadIdx.x * SEG_SCAN_ITEMS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // prefix over earlier tiles (independent of this tile's own data: issue first)
    unsigned long long part = 0;
    for (int j = threadIdx.x; j < tile; j += SEG_SCAN_THREADS) part += s.status[j];

    int32_t c[SEG_SCAN_ITEMS];
#pragma unroll
    for (int i = 0; i < SEG_SCAN_ITEMS / 4; ++i) {
        const int4 v = *reinterpret_cast<const int4*>(s.cnt + base + 4 * i);
        c[4 * i] = v.x; c[4 * i + 1] = v.y; c[4 * i + 2] = v.z; c[4 * i + 3] = v.w;
    }
    uint32_t tsum = 0, tnz = 0;
#pragma unroll
    for (int i = 0; i < SEG_SCAN_ITEMS; ++i) { tsum += c[i]; tnz += c[i] != 0; }

    uint32_t isum = tsum, inz = tnz;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, isum, o);
        const uint32_t b = __shfl_up_sync(0xffffffffu, inz, o);
        if (lane >= o) { isum += a; inz += b; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 31) { sh_wsum[warp] = isum; sh_wnz[warp] = inz; }
    if (lane == 0) sh_part[warp] = part;
    __syncthreads();
    uint32_t wsum = 0, wnz = 0;
#pragma unroll
    for (int w = 0; w < SEG_SCAN_THREADS / 32; ++w)
        if (w < warp) { wsum += sh_wsum[w]; wnz += sh_wnz[w]; }
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
#pragma unroll
        for (int w = 0; w < SEG_SCAN_THREADS / 32; ++w) tot += sh_part[w];
        sh_prefix_sum = static_cast<uint32_t>(tot & 0xffffffffull);
        sh_prefix_nz = static_cast<uint32_t>(tot >> 32);
    }
    __syncthreads();
    uint32_t run = sh_prefix_sum + wsum + isum - tsum;
    uint32_t seg = sh_prefix_nz + wnz + inz - tnz;

#pragma unroll
    for (int i = 0; i < SEG_SCAN_ITEMS; ++i) {
        const int64_t row = base + i;
        if (row == RA) s.totals[2] = static_cast<int32_t>(seg);
        if (c[i] != 0) {
            s.off[row] = static_cast<int32_t>(run);
            if (s.sid) s.sid[row] = static_cast<int32_t>(seg);
            s.seg_row[seg] = static_cast<int32_t>(row);
            s.seg_start[seg] = static_cast<int32_t>(run);
            if (s.long_cap > 0 && c[i] > s.long_cap) s.long_list[atomicAdd(s.totals + 3, 1)] = static_cast<int32_t>(seg);
            run += c[i];
            ++seg;
        }
    }
    if (tile == s.ntiles - 1 && threadIdx.x == SEG_SCAN_THREADS - 1) {
        s.totals[0] = static_cast<int32_t>(seg);
        s.totals[1] = static_cast<int32_t>(run);
        s.seg_start[seg] = static_cast<int32_t>(run);
        if (RA >= s.Rpad) s.totals[2] = static_cast<int32_t>(seg);
    }
}

// Sorts the member lists of the hot rows in place (ascending term id), so the
// consumers can walk them in order.  Term ids are distinct and < Tmax, so the
// rank of a member is the number of set bits below it in a bitmap of the
// segment: bitmap -> per-word prefix popcount (block scan) -> rank.  O(Tmax / 32 +
// len) per hot row, deterministic.  totals[3] is reset by the consumer chain.
static __global__ void __launch_bounds__(256) seg_sort_long_kernel(SegIndex s) {
    __shared__ uint32_t sh_scan[256];
    const int nlong = s.totals[3];
    uint32_t* bits = s.long_bits + static_cast<size_t>(blockIdx.x) * 2 * s.long_words;
    uint32_t* pre = bits + s.long_words;
    const int W = static_cast<int>(s.long_words);
    const int per = (W + 255) / 256;
    for (int li = blockIdx.x; li < nlong; li += gridDim.x) {
        const int seg = s.long_list[li];
        const int start = s.seg_start[seg];
        const int len = s.seg_start[seg + 1] - start;
        for (int w = threadIdx.x; w < W; w += 256) bits[w] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += 256) {
            const int t = s.members[start + i];
            atomicOr(bits + (t >> 5), 1u << (t & 31));
        }
        __syncthreads();
        // exclusive prefix of popcounts: thread owns words [lo, hi)
        const int lo = threadIdx.x * per, hi = lo + per < W ? lo + per : W;
        uint32_t sum = 0;
        for (int w = lo; w < hi; ++w) sum += __popc(bits[w]);
        sh_scan[threadIdx.x] = sum;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const uint32_t v = threadIdx.x >= o ? sh_scan[threadIdx.x - o] : 0u;
            __syncthreads();
            sh_scan[threadIdx.x] += v;
            __syncthreads();
        }
        uint32_t run = sh_scan[threadIdx.x] - sum;
        for (int w = lo; w < hi; ++w) { pre[w] = run; run += __popc(bits[w]); }
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += 256) {
            const int t = s.members[start + i];
            const uint32_t r = pre[t >> 5] + __popc(bits[t >> 5] & ((1u << (t & 31)) - 1u));
            s.long_tmp[start + r] = t;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < len; i += 256) s.members[start + i] = s.long_tmp[start + i];
        __syncthreads();
    }
}

// Host helper: both launches of the scan.
static inline void seg_scan_launch(const SegIndex& s, int64_t RA, cudaStream_t st) {
    seg_tilesum_kernel<<<static_cast<unsigned>(s.ntiles), SEG_SCAN_THREADS, 0, st>>>(s);
    seg_scan_kernel<<<static_cast<unsigned>(s.ntiles), SEG_SCAN_THREADS, 0, st>>>(s, RA);
}

// Re-arms the scan for the next use; run by any later kernel of the chain.
__device__ __forceinline__ void seg_rearm(const SegIndex& s) {
    const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
    (void)nth;
    if (tid == 0 && s.long_cap == 0) s.totals[3] = 0;   // (with hot-row sorting on, the consumer resets it)
}

// members[off[key] + (--cnt[key])] = term.  Afterwards cnt is all-zero again.
__device__ __forceinline__ void seg_place(const SegIndex& s, int64_t key, int32_t term) {
    const int32_t old = atomicSub(s.cnt + key, 1);
    s.members[s.off[key] + old - 1] = term;
}

// Hands `visit(term)` the member terms of one segment in ascending term order.
// G lanes of one group cooperate; all G lanes call with identical (start, len).
// sh: 2 * seg_sort_cap(G) ints of shared scratch private to the group.
__host__ __device__ constexpr int seg_sort_cap(int G) { return G >= 8 ? 128 : (G >= 4 ? 64 : 16 * G); }

template <int G, typename F>
__device__ __forceinline__ void seg_visit_sorted(const int32_t* __restrict__ members, int start,
                                                 int len, int gl /* lane in group */,
                                                 unsigned gmask, int32_t* sh, F visit,
                                                 bool long_presorted = false) {
    if (len == 1) {
        visit(members[start]);
        return;
    }
    if (len == 2) {
        const int32_t a = members[start], b = members[start + 1];
        visit(a < b ? a : b);
        visit(a < b ? b : a);
        return;
    }
    constexpr int CAP = seg_sort_cap(G);
    if (len <= CAP) {
        // rank-by-counting through shared scratch (term ids are distinct)
        int32_t* in = sh;
        int32_t* out = sh + CAP;
        for (int i = gl; i < len; i += G) in[i] = members[start + i];
        __syncwarp(gmask);
        for (int i = gl; i < len; i += G) {
            const int32_t m = in[i];
            int r = 0;
            for (int j = 0; j < len; ++j) r += in[j] < m;
            out[r] = m;
        }
        __syncwarp(gmask);
        for (int i = 0; i < len; ++i) visit(out[i]);
        __syncwarp(gmask);
        return;
    }
    if (long_presorted) {            // hot row already sorted by seg_sort_long_kernel
        for (int i = 0; i < len; ++i) visit(members[start + i]);
        return;
    }
    // long segment without the pre-sort: repeated min-selection straight from
    // global memory.  O(len^2 / G) but correct for any length.
    int32_t last = -1;
    for (int i = 0; i < len; ++i) {
        int32_t best = 0x7fffffff;
        for (int j = gl; j < len; j += G) {
            const int32_t m = members[start + j];
            if (m > last && m < best) best = m;
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const int32_t other = __shfl_xor_sync(gmask, best, o);
            best = other < best ? other : best;
        }
        visit(best);
        last = best;
    }
}

#endif  // __CUDACC__
