// Device-side NumPy-legacy RandomState stream: MT19937 block generation and the
// masked-rejection bounded draw that spotlight/sampling.py:34
// (random_state.randint(0, num_items, shape, dtype=int64)) consumes.
//
// The stream is stored as consecutive 624-word *untempered* key blocks so the
// host can hand the state back to numpy (RandomState.set_state) at any word.
//
// mt19937_fill_kernel: the MT recurrence has dependency distance 227
// (x[k+624] needs x[k], x[k+1], x[k+397]); one CTA advances a block in three
// barrier-separated rounds (227 + 227 + 170 words) out of shared memory and
// streams the blocks to HBM.  It is latency-bound on one SM by construction
// (~10 G words/s); it runs on a side stream ahead of the consumer.
//
// sample_* kernels: parallel stream compaction of the accepted words
// (tile counts -> scan -> ordered scatter), bit-exact with numpy's sequential
// loop because acceptance of a word does not depend on earlier words.
#include "common.cuh"

namespace {

constexpr int MT_N = 624;
constexpr int MT_M = 397;

__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// Launched with ~200 KB of (unused) dynamic shared memory so the CTA owns its SM:
// it runs on a side stream next to the training kernels, and sharing issue slots
// with their CTAs slows this latency-bound loop by an order of magnitude.
__global__ void __launch_bounds__(256) mt19937_fill_kernel(uint32_t* blocks, int64_t nblocks) {
    extern __shared__ uint32_t sm_reserve[];
    __shared__ uint32_t buf[2][MT_N + 1];
    if (nblocks < 0) sm_reserve[0] = 0;     // keep the reservation referenced
    const int t = threadIdx.x;
    for (int k = t; k < MT_N; k += 256) buf[0][k] = blocks[k];
    __syncthreads();
    int cur = 0;
    for (int64_t blk = 1; blk < nblocks; ++blk) {
        const uint32_t* o = buf[cur];
        uint32_t* n = buf[cur ^ 1];
        if (t < 227) n[t] = mt_mix(o[t], o[t + 1], o[t + MT_M]);
        __syncthreads();
        if (t < 227) n[227 + t] = mt_mix(o[227 + t], o[228 + t], n[t]);
        __syncthreads();
        if (t < 169) n[454 + t] = mt_mix(o[454 + t], o[455 + t], n[227 + t]);
        if (t == 169) n[623] = mt_mix(o[623], n[0], n[396]);
        __syncthreads();
        uint32_t* dst = blocks + blk * MT_N;
        for (int k = t; k < MT_N; k += 256) dst[k] = n[k];
        // no 4th barrier: the next round only writes the *old* buffer, whose last
        // readers ran before the 3rd barrier; the stores above read the new one
        cur ^= 1;
    }
}

// Parallel variant: CTA r starts from states[r] (the generator state r*J blocks
// after block 0, produced by the jump kernel; r = 0 is block 0 itself) and
// writes blocks r*J + 1 .. r*J + J.  Only twisted blocks are written, so every
// stored word is exact (a jumped state may carry garbage in the 31 low bits of
// its word 0, which are not part of the MT19937 state).
__global__ void __launch_bounds__(256)
mt19937_fill_par_kernel(uint32_t* blocks, const uint32_t* __restrict__ states, int64_t J, int64_t nblocks) {
    __shared__ uint32_t buf[2][MT_N + 1];
    const int t = threadIdx.x;
    const int64_t r = blockIdx.x;
    const uint32_t* src = r == 0 ? blocks : states + r * MT_N;
    for (int k = t; k < MT_N; k += 256) buf[0][k] = src[k];
    __syncthreads();
    int cur = 0;
    for (int64_t j = 1; j <= J; ++j) {
        const int64_t blk = r * J + j;
        if (blk >= nblocks) break;
        const uint32_t* o = buf[cur];
        uint32_t* n = buf[cur ^ 1];
        if (t < 227) n[t] = mt_mix(o[t], o[t + 1], o[t + MT_M]);
        __syncthreads();
        if (t < 227) n[227 + t] = mt_mix(o[227 + t], o[228 + t], n[t]);
        __syncthreads();
        if (t < 169) n[454 + t] = mt_mix(o[454 + t], o[455 + t], n[227 + t]);
        if (t == 169) n[623] = mt_mix(o[623], n[0], n[396]);
        __syncthreads();
        uint32_t* dst = blocks + blk * MT_N;
        for (int k = t; k < MT_N; k += 256) dst[k] = n[k];
        cur ^= 1;
    }
}

// Jump ahead: out = g(T) in, T = one-word MT19937 transition, g = x^(624 * 2^k) mod
// the minimal polynomial (table from spotlight_b200/data/gen_mt19937_jump.py).
// Horner over the 19968 coefficient bits, 32 bits per iteration: acc <- T^32(acc)
// (32 new words are independent: recurrence distance 227), then
// acc[t] ^= XOR_{m : bit m set} E[t + m] with E = the input state extended by 32
// words, i.e. T^m(in)[t] = E[t + m].
// Round `m` of the doubling schedule: CTA c jumps states[c << (m+1)] by 2^(k+m)
// blocks into states[(c << (m+1)) + (1 << m)].
constexpr int JUMP_THREADS = 640;

// acc <- g(T) in, g given as 624 coefficient words; result written to out[0..624).
__device__ __forceinline__ void mt_jump_apply(const uint32_t* __restrict__ in, const uint32_t* __restrict__ poly,
                                              uint32_t* __restrict__ out, uint32_t* E, uint32_t* acc) {
    const int t = threadIdx.x;
    if (t < MT_N) { E[t] = in[t]; acc[t] = 0u; }
    __syncthreads();
    if (t < 32) E[MT_N + t] = mt_mix(E[t], E[t + 1], E[t + MT_M]);
    __syncthreads();
    // thread t only ever reads E[t .. t + 31] (the same window for every coefficient word): keep
    // it in registers, so the XOR phase of an iteration is 32 register operations and no
    // shared-memory traffic (the loop was bound by 32 predicated shared loads per thread)
    uint32_t e[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) e[b] = t < MT_N ? E[t + b] : 0u;
    int o = 0;                                  // circular base of acc
    for (int w = MT_N - 1; w >= 0; --w) {
        const uint32_t cw = __ldg(poly + w);
        uint32_t nw = 0;
        if (t < 32) {
            int i0 = o + t; if (i0 >= MT_N) i0 -= MT_N;
            int i1 = o + t + 1; if (i1 >= MT_N) i1 -= MT_N;
            int im = o + t + MT_M; if (im >= MT_N) im -= MT_N;
            nw = mt_mix(acc[i0], acc[i1], acc[im]);
        }
        __syncthreads();
        if (t < 32) { int i0 = o + t; if (i0 >= MT_N) i0 -= MT_N; acc[i0] = nw; }
        o += 32; if (o >= MT_N) o -= MT_N;
        __syncthreads();
        if (cw != 0u && t < MT_N) {
            uint32_t v = 0;
#pragma unroll
            for (int b = 0; b < 32; ++b) v ^= e[b] & (0u - ((cw >> b) & 1u));
            int i = o + t; if (i >= MT_N) i -= MT_N;
            acc[i] ^= v;
        }
        __syncthreads();
    }
    if (t < MT_N) { int i = o + t; if (i >= MT_N) i -= MT_N; out[t] = acc[i]; }
}

// Doubling round m over state slots that are `scale` apart: CTA c jumps slot
// (c << (m+1)) * scale by the polynomial handed in into slot that + (1 << m) * scale.
__global__ void __launch_bounds__(JUMP_THREADS)
mt19937_jump_kernel(uint32_t* states, const uint32_t* __restrict__ blocks0,
                    const uint32_t* __restrict__ poly, int m, int P, int scale) {
    __shared__ uint32_t E[MT_N + 32];
    __shared__ uint32_t acc[MT_N];
    const int64_t src_r = static_cast<int64_t>(blockIdx.x << (m + 1)) * scale;
    const int64_t dst_r = src_r + static_cast<int64_t>(1 << m) * scale;
    if (dst_r >= P) return;
    const uint32_t* in = src_r == 0 ? blocks0 : states + src_r * MT_N;
    mt_jump_apply(in, poly, states + dst_r * MT_N, E, acc);
}

// Direct round: slot c * R + r (r = 1..R-1) = slot c * R jumped by r * J0 blocks, with the
// precomputed polynomial x^(624 * J0 * r) (row r - 1 of the direct table).  One launch,
// every CTA independent.
__global__ void __launch_bounds__(JUMP_THREADS)
mt19937_jump_direct_kernel(uint32_t* states, const uint32_t* __restrict__ blocks0,
                           const uint32_t* __restrict__ direct, int R, int P) {
    __shared__ uint32_t E[MT_N + 32];
    __shared__ uint32_t acc[MT_N];
    const int c = blockIdx.x / (R - 1), r = blockIdx.x % (R - 1) + 1;
    const int64_t src_r = static_cast<int64_t>(c) * R;
    const int64_t dst_r = src_r + r;
    if (dst_r >= P) return;
    const uint32_t* in = src_r == 0 ? blocks0 : states + src_r * MT_N;
    mt_jump_apply(in, direct + static_cast<int64_t>(r - 1) * MT_N, states + dst_r * MT_N, E, acc);
}

constexpr int SMP_THREADS = 256;
constexpr int SMP_ITEMS = 8;
constexpr int SMP_TILE = SMP_THREADS * SMP_ITEMS;

__global__ void __launch_bounds__(SMP_THREADS)
sample_count_kernel(const uint32_t* __restrict__ blocks, int64_t nwords, const int64_t* cursor,
                    uint32_t rng, uint32_t mask, uint32_t* tile_cnt) {
    __shared__ uint32_t sh[SMP_THREADS / 32];
    const int64_t start = cursor[0];
    const int64_t base = start + static_cast<int64_t>(blockIdx.x) * SMP_TILE;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < SMP_ITEMS; ++i) {
        const int64_t w = base + i * SMP_THREADS + threadIdx.x;
        if (w < nwords) c += (mt_temper(blocks[w]) & mask) <= rng;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int k = 0; k < SMP_THREADS / 32; ++k) s += sh[k];
        tile_cnt[blockIdx.x] = s;
    }
}

// exclusive scan of tile counts by one block (ntiles is at most a few 10^4)
__global__ void __launch_bounds__(1024)
sample_scan_kernel(const uint32_t* tile_cnt, int64_t ntiles, int64_t* tile_off) {
    __shared__ int64_t sh[1024];
    const int t = threadIdx.x;
    const int64_t per = (ntiles + 1023) / 1024;
    const int64_t lo = t * per, hi = lo + per < ntiles ? lo + per : ntiles;
    int64_t s = 0;
    for (int64_t k = lo; k < hi; ++k) s += tile_cnt[k];
    sh[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int64_t v = t >= o ? sh[t - o] : 0;
        __syncthreads();
        sh[t] += v;
        __syncthreads();
    }
    int64_t run = sh[t] - s;
    for (int64_t k = lo; k < hi; ++k) { tile_off[k] = run; run += tile_cnt[k]; }
    if (t == 1023) tile_off[ntiles] = sh[1023];
}

__global__ void __launch_bounds__(SMP_THREADS)
sample_scatter_kernel(const uint32_t* __restrict__ blocks, int64_t nwords, int64_t* cursor,
                      uint32_t rng, uint32_t mask, int64_t count, const int64_t* tile_off,
                      int64_t ntiles, int64_t* out, int64_t* result) {
    __shared__ uint32_t sh[SMP_THREADS / 32];
    const int64_t start = cursor[0];
    // blocked arrangement so that ranks follow stream order
    const int64_t base = start + static_cast<int64_t>(blockIdx.x) * SMP_TILE + threadIdx.x * SMP_ITEMS;
    uint32_t v[SMP_ITEMS];
    bool ok[SMP_ITEMS];
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < SMP_ITEMS; ++i) {
        const int64_t w = base + i;
        v[i] = w < nwords ? mt_temper(blocks[w]) & mask : 0xffffffffu;
        ok[i] = w < nwords && v[i] <= rng;
        c += ok[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    uint32_t wpre = 0;
    for (int k = 0; k < warp; ++k) wpre += sh[k];
    int64_t rank = tile_off[blockIdx.x] + wpre + inc - c;
#pragma unroll
    for (int i = 0; i < SMP_ITEMS; ++i) {
        if (ok[i]) {
            if (rank < count) out[rank] = static_cast<int64_t>(v[i]);
            if (rank == count - 1) result[0] = base + i + 1;   // one past last consumed word
            ++rank;
        }
    }
    if (blockIdx.x == ntiles - 1 && threadIdx.x == SMP_THREADS - 1) {
        const int64_t total = tile_off[ntiles];
        result[1] = total < count ? total : count;
        if (total < count) result[0] = nwords;                  // stream exhausted
    }
}

__global__ void sample_commit_kernel(int64_t* cursor, const int64_t* result) {
    cursor[0] = result[0];
    cursor[1] = result[1];
}

// Commit + hand-over on the device: the block holding the next unread word becomes block 0
// and the cursor its position in it, exactly what RandomState.set_state would be given
// (numpy leaves pos = 624 on a block boundary).  cursor[2] counts draws that ran out of
// stream words, cursor[3] the values produced since the stream was opened.
__global__ void __launch_bounds__(640)
sample_rebase_kernel(uint32_t* blocks, int64_t nwords, int64_t* cursor, const int64_t* result, int64_t count) {
    __shared__ uint32_t tmp[MT_N];
    const int64_t end = result[0], produced = result[1];
    int64_t blk, pos;
    if (end >= nwords) { blk = nwords / MT_N - 1; pos = MT_N; }
    else if (end % MT_N == 0 && end > 0) { blk = end / MT_N - 1; pos = MT_N; }
    else { blk = end / MT_N; pos = end % MT_N; }
    const int t = threadIdx.x;
    if (t < MT_N) tmp[t] = blocks[blk * MT_N + t];
    __syncthreads();
    if (t < MT_N) blocks[t] = tmp[t];
    if (t == 0) {
        cursor[0] = pos;
        cursor[1] = produced;
        if (produced < count) cursor[2] += 1;
        cursor[3] += produced;
    }
}

}  // namespace

extern "C" {

int slb_mt19937_fill(uint32_t* blocks, int64_t nblocks, slb_stream_t stream) {
    SLB_REQUIRE(blocks != nullptr && nblocks >= 1, "mt19937_fill: bad arguments");
    if (nblocks == 1) return SLB_OK;
    constexpr int kReserve = 200 * 1024;
    static thread_local bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(mt19937_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kReserve) != cudaSuccess) {
            slb_set_error("mt19937_fill: cannot reserve shared memory");
            return SLB_ECUDA;
        }
        configured = true;
    }
    mt19937_fill_kernel<<<1, 256, kReserve, static_cast<cudaStream_t>(stream)>>>(blocks, nblocks);
    SLB_LAUNCH_CHECK("mt19937_fill_kernel");
    return SLB_OK;
}

int slb_mt19937_fill_parallel(uint32_t* blocks, int64_t nblocks, const uint32_t* jump_table,
                              int32_t table_rows, uint32_t* states /* [128 * 624] */,
                              slb_stream_t stream) {
    SLB_REQUIRE(blocks && jump_table && states && nblocks >= 1, "mt19937_fill_parallel: bad arguments");
    if (nblocks == 1) return SLB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // stride J = 2^k blocks per CTA: at most 128 CTAs, and at least 1024 blocks per
    // CTA (a jump round costs about as much as generating ~1000 blocks)
    int k = 10;
    while (((nblocks - 1 + (1ll << k) - 1) >> k) > 128) ++k;
    const int64_t J = 1ll << k;
    const int P = static_cast<int>((nblocks - 1 + J - 1) / J);
    int top = 0;
    while ((1 << (top + 1)) < P) ++top;               // highest m with 2^m < P
    SLB_REQUIRE(k + top < table_rows, "mt19937_fill_parallel: jump table too small for %lld blocks",
                static_cast<long long>(nblocks));
    if (P > 1) {
        for (int m = top; m >= 0; --m) {
            const int grid = (P - (1 << m) + (1 << (m + 1)) - 1) >> (m + 1);
            if (grid <= 0) continue;
            mt19937_jump_kernel<<<grid, JUMP_THREADS, 0, st>>>(states, blocks,
                                                               jump_table + static_cast<int64_t>(k + m) * MT_N, m, P, 1);
            SLB_LAUNCH_CHECK("mt19937_jump_kernel");
        }
    }
    mt19937_fill_par_kernel<<<P, 256, 0, st>>>(blocks, states, J, nblocks);
    SLB_LAUNCH_CHECK("mt19937_fill_par_kernel");
    return SLB_OK;
}

int64_t slb_mt19937_direct_slots(int64_t nblocks, int32_t j0_log2) {
    const int64_t J0 = 1ll << j0_log2;
    return nblocks <= 1 ? 1 : (nblocks - 1 + J0 - 1) / J0;
}

int slb_mt19937_fill_direct(uint32_t* blocks, int64_t nblocks, const uint32_t* jump_table,
                            int32_t table_rows, const uint32_t* direct_table, int32_t direct_rows,
                            int32_t j0_log2, uint32_t* states, int64_t state_slots, slb_stream_t stream) {
    SLB_REQUIRE(blocks && jump_table && direct_table && states && nblocks >= 1, "mt19937_fill_direct: bad arguments");
    SLB_REQUIRE(direct_rows >= 1 && j0_log2 >= 0 && j0_log2 < 20, "mt19937_fill_direct: bad table shape");
    if (nblocks == 1) return SLB_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t J0 = 1ll << j0_log2;
    const int R = direct_rows + 1;                                   // fine slots per coarse slot
    const int64_t P64 = (nblocks - 1 + J0 - 1) / J0;                 // fine slots (CTAs of the fill)
    SLB_REQUIRE(P64 <= state_slots && P64 < (1ll << 30), "mt19937_fill_direct: %lld state slots needed, %lld given",
                static_cast<long long>(P64), static_cast<long long>(state_slots));
    const int P = static_cast<int>(P64);
    const int C = (P + R - 1) / R;                                   // coarse slots
    if (C > 1) {
        // coarse slots are R * J0 blocks apart: doubling rounds with the power-of-two table
        int kc = j0_log2;
        while ((1 << (kc - j0_log2)) < R) ++kc;
        SLB_REQUIRE((1 << (kc - j0_log2)) == R, "mt19937_fill_direct: direct_rows + 1 must be a power of two");
        int top = 0;
        while ((1 << (top + 1)) < C) ++top;
        SLB_REQUIRE(kc + top < table_rows, "mt19937_fill_direct: jump table too small for %lld blocks",
                    static_cast<long long>(nblocks));
        for (int m = top; m >= 0; --m) {
            const int grid = (C - (1 << m) + (1 << (m + 1)) - 1) >> (m + 1);
            if (grid <= 0) continue;
            mt19937_jump_kernel<<<grid, JUMP_THREADS, 0, st>>>(states, blocks,
                                                               jump_table + static_cast<int64_t>(kc + m) * MT_N, m, P, R);
            SLB_LAUNCH_CHECK("mt19937_jump_kernel");
        }
    }
    if (P > 1) {
        mt19937_jump_direct_kernel<<<C * (R - 1), JUMP_THREADS, 0, st>>>(states, blocks, direct_table, R, P);
        SLB_LAUNCH_CHECK("mt19937_jump_direct_kernel");
    }
    mt19937_fill_par_kernel<<<P, 256, 0, st>>>(blocks, states, J0, nblocks);
    SLB_LAUNCH_CHECK("mt19937_fill_par_kernel");
    return SLB_OK;
}

size_t slb_sample_workspace_bytes(int64_t nwords) {
    const int64_t ntiles = (nwords + SMP_TILE - 1) / SMP_TILE + 1;
    WsCarver ws(nullptr);
    ws.take<uint32_t>(ntiles);
    ws.take<int64_t>(ntiles + 1);
    ws.take<int64_t>(2);
    return ws.bytes();
}

static int sample_bounded_impl(uint32_t* blocks, int64_t nwords, int64_t* cursor, uint32_t rng,
                               int64_t count, int64_t* out, void* workspace, size_t workspace_bytes,
                               bool chain, slb_stream_t stream) {
    SLB_REQUIRE(blocks && cursor && out && workspace, "sample_bounded: null pointer");
    SLB_REQUIRE(count > 0 && nwords > 0, "sample_bounded: count and nwords must be > 0");
    SLB_REQUIRE(rng != 0 && rng != 0xffffffffu, "sample_bounded: rng must be in [1, 2^32 - 2]");
    SLB_REQUIRE(!chain || nwords % MT_N == 0, "sample_bounded_chain: nwords must be whole blocks");
    if (workspace_bytes < slb_sample_workspace_bytes(nwords)) {
        slb_set_error("sample_bounded: workspace too small");
        return SLB_ENOSPC;
    }
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    // tiles are laid out from the cursor; the host bounds them by nwords (cursor >= 0)
    const int64_t ntiles = (nwords + SMP_TILE - 1) / SMP_TILE;
    WsCarver ws(workspace);
    uint32_t* tile_cnt = ws.take<uint32_t>(ntiles + 1);
    int64_t* tile_off = ws.take<int64_t>(ntiles + 1);
    int64_t* result = ws.take<int64_t>(2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sample_count_kernel<<<static_cast<unsigned>(ntiles), SMP_THREADS, 0, st>>>(blocks, nwords, cursor, rng, mask, tile_cnt);
    SLB_LAUNCH_CHECK("sample_count_kernel");
    sample_scan_kernel<<<1, 1024, 0, st>>>(tile_cnt, ntiles, tile_off);
    SLB_LAUNCH_CHECK("sample_scan_kernel");
    sample_scatter_kernel<<<static_cast<unsigned>(ntiles), SMP_THREADS, 0, st>>>(blocks, nwords, cursor, rng, mask, count,
                                                                 tile_off, ntiles, out, result);
    SLB_LAUNCH_CHECK("sample_scatter_kernel");
    if (chain) {
        sample_rebase_kernel<<<1, 640, 0, st>>>(blocks, nwords, cursor, result, count);
        SLB_LAUNCH_CHECK("sample_rebase_kernel");
    } else {
        sample_commit_kernel<<<1, 1, 0, st>>>(cursor, result);
        SLB_LAUNCH_CHECK("sample_commit_kernel");
    }
    return SLB_OK;
}

int slb_sample_bounded(const uint32_t* blocks, int64_t nwords, int64_t* cursor, uint32_t rng,
                       int64_t count, int64_t* out, void* workspace, size_t workspace_bytes,
                       slb_stream_t stream) {
    return sample_bounded_impl(const_cast<uint32_t*>(blocks), nwords, cursor, rng, count, out, workspace,
                               workspace_bytes, false, stream);
}

int slb_sample_bounded_chain(uint32_t* blocks, int64_t nwords, int64_t* cursor, uint32_t rng,
                             int64_t count, int64_t* out, void* workspace, size_t workspace_bytes,
                             slb_stream_t stream) {
    return sample_bounded_impl(blocks, nwords, cursor, rng, count, out, workspace, workspace_bytes, true, stream);
}

}  // extern "C"
