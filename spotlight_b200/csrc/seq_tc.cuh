// tcgen05 (5th-generation tensor core) path of the CNNNet causal convolution for
// D = 128 -- BASELINE configs[4].  Included by seq.cu after ConvGemm / ConvDw.
//
// The conv is a dense contraction: per layer
//   forward / input-gradient : Out[(b,t), n] = sum_{j<k} sum_c In[b, t + shift_j, c] * W_j[n][c]
//   weight gradient          : dW_j[i][o]    = sum_{(b,t)} In[b, t + shift_j, i] * dZ[(b,t), o]
// Both run as 128 x 128 output tiles with the accumulator in TMEM (128 lanes x
// 128 fp32 columns) and tcgen05.mma.kind::tf32 issued by one thread.  fp32-level
// accuracy (the 1e-5 parity budget) comes from the 3xTF32 split: every operand
// chunk is staged twice in shared memory (hi = tf32(x), lo = tf32(x - hi)) and
// each K step issues lo*hi + hi*lo + hi*hi into the same accumulator.
//
// Operands are staged by the CUDA cores (the split has to touch every element
// anyway) straight into the no-swizzle canonical UMMA layouts:
//   K-major  (forward/dX): 8 rows x 16 B core matrices, element (r, k) at
//            (r/8)*SBO + (k/4)*LBO + (r%8)*16 + (k%4)*4,  LBO = 128, SBO = 1024
// The weight gradient contracts over positions, which are the *outer* index of its
// operands in memory; its staging transposes 4 x 4 register blocks so it can use the
// same K-major layout.
// One mbarrier tracks MMA completion (tcgen05.commit); the pipeline is single
// stage (stage -> fence -> MMA -> wait), which already moves the math off the
// CUDA cores; multi-stage TMA feeding is the next step.
#pragma once

namespace tc {

constexpr int TM = 128;          // tile rows (positions / in-channels)
constexpr int TN = 128;          // tile cols (= D)
constexpr int KC = 32;           // K elements staged per chunk (4 MMA k-steps of 8)
constexpr int TILE_BYTES = TM * KC * 4;          // 16 KB per staged operand copy
constexpr int SMEM_BYTES = 4 * TILE_BYTES + 64;  // A_hi, A_lo, B_hi, B_lo + barrier + tmem ptr

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a wrong descriptor must fail loudly (trap), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 24); ++spin)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, int ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, int ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu) |
           (static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16) |
           (static_cast<uint64_t>((sbo >> 4) & 0x3fffu) << 32) |
           (1ull << 46);
}
// instruction descriptor: D = f32, A = B = tf32, M = 128, N = 128
constexpr uint32_t IDESC_KK = (1u << 4) | (2u << 7) | (2u << 10) | ((TN >> 3) << 17) | ((TM >> 4) << 24);

__device__ __forceinline__ void split4(float4 x, uint4& hi, uint4& lo) {
    split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y);
    split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
}

// 3 x (KC / 8) MMAs for one staged chunk; `first` clears the accumulator on the first one
__device__ __forceinline__ void issue_chunk(uint32_t tmem, uint32_t sA_hi, uint32_t sA_lo, uint32_t sB_hi,
                                            uint32_t sB_lo, uint32_t step_bytes, uint32_t lbo, uint32_t sbo,
                                            uint32_t idesc, bool first) {
#pragma unroll
    for (int s = 0; s < KC / 8; ++s) {
        const uint64_t ah = make_desc(sA_hi + s * step_bytes, lbo, sbo), al = make_desc(sA_lo + s * step_bytes, lbo, sbo);
        const uint64_t bh = make_desc(sB_hi + s * step_bytes, lbo, sbo), bl = make_desc(sB_lo + s * step_bytes, lbo, sbo);
        umma_tf32(tmem, al, bh, idesc, (first && s == 0) ? 0u : 1u);
        umma_tf32(tmem, ah, bl, idesc, 1u);
        umma_tf32(tmem, ah, bh, idesc, 1u);
    }
}

// ---------------------------------------------------------------------------
// forward / input-gradient:  K-major operands.  g.Wm is [k][n][c] (row n holds the
// contraction index contiguously): Wb for the forward, Wf for the input gradient.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_conv_gemm_kernel(ConvGemm g) {
    extern __shared__ __align__(128) uint8_t tc_smem[];
    uint8_t* A_hi = tc_smem;
    uint8_t* A_lo = tc_smem + TILE_BYTES;
    uint8_t* B_hi = tc_smem + 2 * TILE_BYTES;
    uint8_t* B_lo = tc_smem + 3 * TILE_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tc_smem + 4 * TILE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_smem + 4 * TILE_BYTES + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int D = g.D;                                   // == 128
    const int64_t M = g.B * g.Tout;
    const int64_t m = static_cast<int64_t>(blockIdx.x) * TM + tid;     // this thread's output row
    const int64_t b = m < M ? m / g.Tout : 0;
    const int t = m < M ? static_cast<int>(m - b * g.Tout) : 0;

    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(tmem_slot, TN);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;

    const uint32_t off = (tid >> 3) * 1024 + (tid & 7) * 16;          // (r/8)*SBO + (r%8)*16
    uint32_t phase = 0;
    bool first = true;
    for (int j = 0; j < g.k; ++j) {
        const int q = t + g.shift[j];
        const bool rowok = m < M && q >= 0 && q < g.Tin;
        const float* arow = g.In + (b * g.Tin + (rowok ? q : 0)) * D;
        const float* brow = g.Wm + (static_cast<int64_t>(j) * D + tid) * D;     // row n = tid
        for (int c0 = 0; c0 < D; c0 += KC) {
#pragma unroll
            for (int c = 0; c < KC / 4; ++c) {
                float4 av = make_float4(0, 0, 0, 0);
                if (rowok) av = ld4(arow + c0 + 4 * c);
                const float4 bv = ldg4(brow + c0 + 4 * c);
                uint4 h, l;
                split4(av, h, l);
                *reinterpret_cast<uint4*>(A_hi + off + c * 128) = h;
                *reinterpret_cast<uint4*>(A_lo + off + c * 128) = l;
                split4(bv, h, l);
                *reinterpret_cast<uint4*>(B_hi + off + c * 128) = h;
                *reinterpret_cast<uint4*>(B_lo + off + c * 128) = l;
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                fence_after();
                issue_chunk(tmem, smem_u32(A_hi), smem_u32(A_lo), smem_u32(B_hi), smem_u32(B_lo),
                            256, 128, 1024, IDESC_KK, first);
                umma_commit(bar);
            }
            first = false;
            mbar_wait(bar, phase);           // MMAs done: operands may be overwritten
            phase ^= 1;
        }
    }
    fence_after();

    // epilogue: thread = output row `m`, 4 x 32 columns out of TMEM lane 32*warp + lane
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const int rt = t + g.res_shift;
    const bool resok = g.Res && m < M && rt >= 0 && rt < g.res_T;
    for (int cb = 0; cb < TN; cb += 32) {
        float v[32];
        tmem_ld32(lane_base + cb, v);
        if (m < M) {
#pragma unroll
            for (int c4 = 0; c4 < 32; c4 += 4) {
                const int n = cb + c4;
                float x[4] = {v[c4], v[c4 + 1], v[c4 + 2], v[c4 + 3]};
                float4 res = make_float4(0, 0, 0, 0);
                if (resok) res = ld4(g.Res + (b * g.res_T + rt) * D + n);
                if (g.mode == 0) {
                    const float4 bb = ldg4(g.bias + n);
                    x[0] += bb.x; x[1] += bb.y; x[2] += bb.z; x[3] += bb.w;
#pragma unroll
                    for (int y = 0; y < 4; ++y) x[y] = g.nonlin == 0 ? tanhf(x[y]) : fmaxf(x[y], 0.f);
                    st4(g.Aout + m * D + n, make_float4(x[0], x[1], x[2], x[3]));
                }
                float4 o = make_float4(x[0] + res.x, x[1] + res.y, x[2] + res.z, x[3] + res.w);
                if (g.accumulate) {
                    const float4 old = ld4(g.Out + m * D + n);
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                st4(g.Out + m * D + n, o);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TN);
}

// ---------------------------------------------------------------------------
// weight gradient (positions are the contraction index):
//   A[row = i][k = pos] = In[b, t + shift_j, i],  B[row = o][k = pos] = dZ[pos, o]
// grid = (k taps, splits); every CTA owns the whole 128 x 128 (i, o) tile of one tap.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_conv_dw_kernel(ConvDw g) {
    extern __shared__ __align__(128) uint8_t tc_smem[];
    uint8_t* A_hi = tc_smem;
    uint8_t* A_lo = tc_smem + TILE_BYTES;
    uint8_t* B_hi = tc_smem + 2 * TILE_BYTES;
    uint8_t* B_lo = tc_smem + 3 * TILE_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tc_smem + 4 * TILE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_smem + 4 * TILE_BYTES + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int D = g.D;                                   // == 128
    const int j = blockIdx.x;
    const int64_t split = blockIdx.y;
    const int64_t M = g.B * g.Tout;
    const int64_t mlo = split * g.slab, mhi = mlo + g.slab < M ? mlo + g.slab : M;

    if (tid == 0) mbar_init(bar, 1);
    if (warp == 0) tmem_alloc(tmem_slot, TN);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;

    // staging: both operands are brought into the K-major canonical layout (rows =
    // channel, k = position) by a 4 x 4 register transpose: a unit is 4 positions x 4
    // channels; thread handles units (pg = u / 32, cg = u % 32) for u = tid, tid + 128.
    uint32_t phase = 0;
    bool first = true;
    float bacc = 0.f;                                    // bias gradient: column sums of dZ (tap 0 only)
    for (int64_t mb = mlo; mb < mhi; mb += KC) {
#pragma unroll
        for (int uu = 0; uu < 2; ++uu) {
            const int u = tid + uu * 128;
            const int pg = u >> 5, cg = u & 31;
            float4 a4[4], b4[4];
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const int64_t mm = mb + pg * 4 + pp;
                const bool ok = mm < mhi;
                const int64_t bb = ok ? mm / g.Tout : 0;
                const int t = ok ? static_cast<int>(mm - bb * g.Tout) : 0;
                const int q = t + g.shift[j];
                a4[pp] = make_float4(0, 0, 0, 0);
                b4[pp] = make_float4(0, 0, 0, 0);
                if (ok && q >= 0 && q < g.Tin) a4[pp] = ld4(g.In + (bb * g.Tin + q) * D + cg * 4);
                if (ok) b4[pp] = ld4(g.dZ + mm * D + cg * 4);
            }
            // channel i = 4*cg + q owns the 16-byte chunk holding positions 4*pg .. 4*pg+3
            const float ar[4][4] = {{a4[0].x, a4[1].x, a4[2].x, a4[3].x}, {a4[0].y, a4[1].y, a4[2].y, a4[3].y},
                                    {a4[0].z, a4[1].z, a4[2].z, a4[3].z}, {a4[0].w, a4[1].w, a4[2].w, a4[3].w}};
            const float br[4][4] = {{b4[0].x, b4[1].x, b4[2].x, b4[3].x}, {b4[0].y, b4[1].y, b4[2].y, b4[3].y},
                                    {b4[0].z, b4[1].z, b4[2].z, b4[3].z}, {b4[0].w, b4[1].w, b4[2].w, b4[3].w}};
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const int row = cg * 4 + qd;
                const uint32_t o = (row >> 3) * 1024 + pg * 128 + (row & 7) * 16;
                uint4 h, l;
                split4(make_float4(ar[qd][0], ar[qd][1], ar[qd][2], ar[qd][3]), h, l);
                *reinterpret_cast<uint4*>(A_hi + o) = h;
                *reinterpret_cast<uint4*>(A_lo + o) = l;
                split4(make_float4(br[qd][0], br[qd][1], br[qd][2], br[qd][3]), h, l);
                *reinterpret_cast<uint4*>(B_hi + o) = h;
                *reinterpret_cast<uint4*>(B_lo + o) = l;
            }
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            fence_after();
            issue_chunk(tmem, smem_u32(A_hi), smem_u32(A_lo), smem_u32(B_hi), smem_u32(B_lo),
                        256, 128, 1024, IDESC_KK, first);
            umma_commit(bar);
        }
        first = false;
        if (j == 0) {                                     // db[o = tid]: fixed order over the chunk
            for (int pp = 0; pp < KC; ++pp) {
                const int64_t m2 = mb + pp;
                if (m2 < mhi) bacc += g.dZ[m2 * D + tid];
            }
        }
        mbar_wait(bar, phase);
        phase ^= 1;
    }
    fence_after();
    float* out = g.part + ((split * g.k + j) * D) * D;   // [i][o]
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int cb = 0; cb < TN; cb += 32) {
        float v[32];
        if (mlo < mhi) tmem_ld32(lane_base + cb, v);
        else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int c4 = 0; c4 < 32; c4 += 4)
            st4(out + static_cast<int64_t>(tid) * D + cb + c4, make_float4(v[c4], v[c4 + 1], v[c4 + 2], v[c4 + 3]));
    }
    if (j == 0) g.bpart[split * D + tid] = bacc;
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, TN);
}

}  // namespace tc
