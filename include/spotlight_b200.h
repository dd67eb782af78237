/*
 * spotlight_b200 C-ABI  --  the drop-in boundary of the B200-native
 * implicit-feedback fit() hot path.
 *
 * The reference (maciejkula/spotlight) has no FFI layer: its hot path is
 * Python calling stock ATen ops.  Each entry point below replaces the span of
 * reference Python named in its comment (file:line under the reference root).
 * The Python host (spotlight_b200/_lib.py, ops.py) binds these with ctypes and
 * registers them as torch.library custom ops; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - C99 POD arguments only: raw *device* pointers + explicit sizes.  No
 *     torch types.  Every buffer, including workspaces, is owned by the caller.
 *   - Stateless and re-entrant: no global mutable state except a thread-local
 *     error string.  Never allocates device memory, never synchronises.
 *   - All work is enqueued asynchronously on `stream` (a cudaStream_t passed
 *     as void*).
 *   - Return value: 0 on success, negative SLB_E* on error; message through
 *     slb_last_error().
 *   - ids are int64 (the reference's hot-path id dtype,
 *     spotlight/factorization/implicit.py:202-203); parameters are fp32
 *     row-major [rows, dim].
 */
#ifndef SPOTLIGHT_B200_H
#define SPOTLIGHT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLB_VERSION 100

#define SLB_OK 0
#define SLB_EINVAL (-1)   /* bad argument                       */
#define SLB_ECUDA (-2)    /* CUDA launch / runtime error        */
#define SLB_ENOSPC (-3)   /* caller workspace too small         */

/* loss kinds: spotlight/losses.py:18,53,93,127 */
#define SLB_LOSS_POINTWISE 0
#define SLB_LOSS_BPR 1
#define SLB_LOSS_HINGE 2
#define SLB_LOSS_ADAPTIVE_HINGE 3

/* gradient output modes of the fused training step */
#define SLB_GRAD_DENSE 0    /* write into caller-zeroed dense [rows, dim] grads (reference .grad) */
#define SLB_GRAD_COMPACT 1  /* unique touched rows + their grad rows (sparse=True / fused optimizer) */

/* fused row-wise optimizers (apply only to touched rows) */
#define SLB_OPT_NONE 0
#define SLB_OPT_SGD 1
#define SLB_OPT_ADAGRAD 2

typedef void* slb_stream_t; /* cudaStream_t */

int slb_version(void);
const char* slb_last_error(void);
/* number of SMs of the current device (grid sizing; 148 on B200) */
int slb_sm_count(void);

/* ------------------------------------------------------------------------
 * S1  negative sampling  --  replaces spotlight/sampling.py:31-36
 *     (RandomState.randint(0, num_items, shape, dtype=int64)) and the
 *     MT19937 stream it consumes.
 *
 * The stream lives on the device as consecutive 624-word *untempered* key
 * blocks: block 0 is RandomState.get_state()[1], block j its j-th twist.
 * Word w of the stream is temper(blocks[w]).
 * ---------------------------------------------------------------------- */

/* Fill blocks[1..nblocks) from blocks[0] (each block = one MT19937 twist of
 * the previous one).  blocks: device uint32[nblocks*624]. */
int slb_mt19937_fill(uint32_t* blocks, int64_t nblocks, slb_stream_t stream);

size_t slb_sample_workspace_bytes(int64_t nwords);

/* Legacy masked-rejection draw of `count` values on [0, rng] from stream
 * words [*cursor, nwords): out[k] = k-th word w with (temper(w) & mask) <= rng.
 * cursor: device int64[2]; in: cursor[0] = first unread word; out: cursor[0] =
 * one past the last word consumed, cursor[1] = number of values produced
 * (== count on success; < count means the caller must extend the stream and
 * call again for the remainder). */
int slb_sample_bounded(const uint32_t* blocks, int64_t nwords, int64_t* cursor,
                       uint32_t rng, int64_t count, int64_t* out,
                       void* workspace, size_t workspace_bytes, slb_stream_t stream);

/* ------------------------------------------------------------------------
 * E1/E2/E3  embedding gathers -- replaces ScaledEmbedding / ZeroEmbedding /
 * BloomEmbedding forward (spotlight/layers.py:23-56, 206-244) and
 * aten::embedding_dense_backward for them.
 * ---------------------------------------------------------------------- */

/* out[n, dim] = W[ids[n]]   (hash_count == 0)
 * out[n, dim] = sum_k W[h_k(ids[n])], h_k = murmur3_32(le32(id), seeds[k]) floor-mod
 *   rows, h_k = 0 when id == padding_idx   (hash_count = H > 0; layers.py:178-204) */
int slb_embedding_forward(const float* W, int64_t rows, int32_t dim,
                          const int64_t* ids, int64_t n,
                          int32_t hash_count, const uint32_t* seeds /* host, H */,
                          int64_t padding_idx, float* out, slb_stream_t stream);

/* hashed row ids only: rows_out[n, H] int64 (BloomEmbedding._get_hashed_indices) */
int slb_bloom_rows(const int64_t* ids, int64_t n, int32_t hash_count,
                   const uint32_t* seeds /* host */, int64_t rows, int64_t padding_idx,
                   int64_t* rows_out, slb_stream_t stream);

size_t slb_embedding_backward_workspace_bytes(int64_t n_terms, int64_t rows);

/* Deterministic segmented scatter-add:  dW[r] = sum_{t : row(t) == r} dout[t / fan],
 * summed in ascending t.  One writer per row, no float atomics.
 *   hash_count == 0: row(t) = ids[t], fan = 1
 *   hash_count == H: row(t) = h_{t % H}(ids[t / H]), fan = H
 * dW must be zero-filled by the caller (rows not touched are not written);
 * row `frozen_row` (padding_idx, or -1) is left zero.  workspace must have
 * been zero-initialised once (slb_workspace_init) and is left reusable. */
int slb_embedding_backward(const float* dout, const int64_t* ids, int64_t n,
                           int32_t hash_count, const uint32_t* seeds /* host */,
                           int64_t rows, int32_t dim, int64_t frozen_row,
                           float* dW, void* workspace, size_t workspace_bytes,
                           slb_stream_t stream);

/* one-time zero-initialisation of any workspace handed to this library */
int slb_workspace_init(void* workspace, size_t workspace_bytes, slb_stream_t stream);

/* ------------------------------------------------------------------------
 * N1  BilinearNet.forward -- replaces
 *     spotlight/factorization/representations.py:80-91 (inference / predict)
 * scores[n] = <Wu[users[n]], Wi[items[n]]> + bu[users[n]] + bi[items[n]]
 * users may have stride 0 semantics via user_broadcast != 0 (one user, many items,
 * as spotlight/factorization/_components.py:19-20 expands).
 * ---------------------------------------------------------------------- */
int slb_mf_scores(const float* Wu, const float* Wi, const float* bu, const float* bi,
                  int32_t dim, const int64_t* users, const int64_t* items, int64_t n,
                  int32_t user_broadcast, float* scores, slb_stream_t stream);

/* autograd of the above: dense grads of (Wu, Wi, bu, bi) given d loss / d scores.
 * Deterministic segmented scatter; d* must be caller-zeroed.
 * workspace: slb_mf_step_workspace_bytes((n + 1) / 2, 1, 0, num_users, num_items),
 * zero-initialised once. */
int slb_mf_scores_backward(const float* gscores, const int64_t* users, const int64_t* items,
                           int64_t n, int32_t user_broadcast, const float* Wu, const float* Wi,
                           int64_t num_users, int64_t num_items, int32_t dim,
                           float* dWu, float* dWi, float* dbu, float* dbi,
                           void* workspace, size_t workspace_bytes, slb_stream_t stream);

/* ------------------------------------------------------------------------
 * N1+L1..L4+G1  the fused training step -- replaces the loop body of
 *     ImplicitFactorizationModel.fit, spotlight/factorization/implicit.py:229-242
 *     (two BilinearNet forwards, loss, loss.backward()).
 * ---------------------------------------------------------------------- */
typedef struct slb_mf_step_args {
    /* minibatch */
    int64_t batch;            /* B */
    const int64_t* users;     /* [B] */
    const int64_t* items;     /* [B] */
    const int64_t* negs;      /* [B] or [B*n_neg] flat for adaptive hinge (implicit.py:266-275) */
    int32_t loss;             /* SLB_LOSS_* */
    int32_t n_neg;            /* 1 unless adaptive hinge */
    /* parameters (BilinearNet, representations.py:49-59) */
    int64_t num_users, num_items;
    int32_t dim;
    float* Wu; float* Wi;     /* [num_users, dim], [num_items, dim] */
    float* bu; float* bi;     /* [num_users], [num_items] */
    /* outputs */
    float* loss_out;          /* [1]  mean loss of this minibatch */
    float* pos_out;           /* [B] or NULL */
    float* neg_out;           /* [B*n_neg] ((n_neg, B) view) or NULL */
    int32_t grad_mode;        /* SLB_GRAD_DENSE / SLB_GRAD_COMPACT */
    /* dense mode: caller-zeroed full-size grads */
    float* dWu; float* dWi; float* dbu; float* dbi;
    /* compact mode: caller-provided buffers sized by slb_mf_compact_rows();
     * rows are ascending; counts land in compact_counts[0] (users), [1] (items) */
    int64_t* urows; float* gWu; float* gbu;
    int64_t* irows; float* gWi; float* gbi;
    int32_t* compact_counts;  /* device int32[2] */
    /* fused row-wise optimizer on the compact grads (SLB_OPT_NONE to skip) */
    int32_t opt;
    float lr;
    float weight_decay;       /* added as wd*W[row] on touched rows */
    float eps;                /* adagrad */
    float* state_Wu; float* state_Wi; float* state_bu; float* state_bi; /* adagrad sums */
    /* multi-GPU hooks (0 = single-GPU behaviour):
     *  norm_batch   > 0: loss and gradients are normalised by this (global) batch
     *                    size instead of `batch`; loss_out then holds this rank's
     *                    share of the global mean.
     *  opt_users_only != 0: with opt != NONE the optimizer is fused only into the
     *                    user rows; item gradients are written per grad_mode and
     *                    left to the caller (they belong to other ranks' shards). */
    int64_t norm_batch;
    int32_t opt_users_only;
    /* workspace */
    void* workspace; size_t workspace_bytes;
} slb_mf_step_args;

size_t slb_mf_step_workspace_bytes(int64_t batch, int32_t n_neg, int32_t loss,
                                   int64_t num_users, int64_t num_items);
/* upper bound on touched user rows / item rows for compact buffers */
int64_t slb_mf_compact_rows(int64_t batch, int32_t n_neg, int32_t loss, int32_t which /*0 users,1 items*/);

int slb_mf_train_step(const slb_mf_step_args* args, slb_stream_t stream);

/* Profiling aid: launch only the kernels selected by `phases`
 * (1 forward, 2 index scan, 4 index fill, 8 backward, 16 optimizer); calling
 * it once per bit, in that order, equals slb_mf_train_step.  bench.py uses it
 * to time each kernel with CUDA events. */
int slb_mf_train_step_phases(const slb_mf_step_args* args, int32_t phases, slb_stream_t stream);

/* M1  epoch pipeline: runs ceil(n / batch) consecutive training steps (last
 * one short, torch_utils.py:22-32) over device-resident shuffled ids with the
 * fused optimizer, without returning to the host between steps.
 * losses_out: device float[ceil(n / batch)].  `step` carries every per-step
 * field; users/items/negs/batch/loss_out are overridden per step. */
int slb_mf_fit_epoch(const slb_mf_step_args* step, const int64_t* users, const int64_t* items,
                     const int64_t* negs, int64_t n, float* losses_out, slb_stream_t stream);

/* ------------------------------------------------------------------------
 * (e) multi-GPU routing: per-batch index bucketing for range-sharded item rows.
 * Given n ids in [0, rows): uniq = the distinct ids ascending (so grouped by
 * owner = id / chunk), inverse[t] = position of ids[t] in uniq, bounds[p] =
 * first position in uniq owned by rank p (p = 0..nparts; bounds[nparts] = count).
 * counts_out: device int64[nparts + 2] = bounds followed by the unique count.
 * ---------------------------------------------------------------------- */
size_t slb_unique_workspace_bytes(int64_t n, int64_t rows);
int slb_unique_bucket(const int64_t* ids, int64_t n, int64_t rows, int64_t chunk, int32_t nparts,
                      int64_t* uniq /* [min(n, rows)] */, int64_t* inverse /* [n] */,
                      int64_t* counts_out, void* workspace, size_t workspace_bytes,
                      slb_stream_t stream);

/* ------------------------------------------------------------------------
 * L1..L4 standalone losses -- replaces spotlight/losses.py:18-166 for the
 * generic (custom representation) path.  neg is [n_neg, n] for adaptive.
 * mask may be NULL (plain mean).  gpos/gneg may be NULL (forward only).
 * ---------------------------------------------------------------------- */
size_t slb_loss_workspace_bytes(int64_t n);
int slb_pairwise_loss(int32_t loss, const float* pos, const float* neg, const uint8_t* mask,
                      int64_t n, int32_t n_neg, float* loss_out, float* gpos, float* gneg,
                      void* workspace, size_t workspace_bytes, slb_stream_t stream);

/* ------------------------------------------------------------------------
 * Q1  PoolNet -- replaces spotlight/sequence/representations.py:91-114,136-144
 * and the training step of spotlight/sequence/implicit.py:230-255.
 * ---------------------------------------------------------------------- */
typedef struct slb_seq_step_args {
    int64_t batch;            /* B sequences */
    int32_t seq_len;          /* S */
    const int64_t* seqs;      /* [B, S], 0 = padding */
    const int64_t* negs;      /* [B, S] or [n_neg*B, S] (adaptive) */
    int32_t loss; int32_t n_neg;
    int64_t num_items; int32_t dim;
    float* E; float* bias;    /* [num_items, dim], [num_items] */
    /* CNNNet only (representations.py:357-368); n_layers == 0 -> PoolNet */
    int32_t n_layers;
    const int32_t* kernel_width; const int32_t* dilation; /* host arrays [n_layers] */
    int32_t nonlinearity;     /* 0 tanh, 1 relu */
    int32_t residual;
    float* const* conv_w;     /* host array of device ptrs [n_layers], each (D, D, k, 1) */
    float* const* conv_b;     /* host array of device ptrs [n_layers], each (D) */
    float* const* dconv_w;    /* grads, accumulated into caller-zeroed buffers */
    float* const* dconv_b;
    /* outputs */
    float* loss_out; float* pos_out; float* neg_out;
    float* dE; float* dbias;  /* caller-zeroed dense grads */
    void* workspace; size_t workspace_bytes;
} slb_seq_step_args;

size_t slb_seq_step_workspace_bytes(const slb_seq_step_args* args);
int slb_seq_train_step(const slb_seq_step_args* args, slb_stream_t stream);

/* user_representation only (predict path): rep_out [B, S+1, D] time-major */
int slb_seq_representation(const slb_seq_step_args* args, float* rep_out, slb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPOTLIGHT_B200_H */
