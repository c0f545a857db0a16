/* moma_b200 -- C ABI of the B200 (sm_100a) MoMA criterion hot path.
 *
 * The reference (trinhvg/MoMA) has no FFI: its operator interface is the Python
 * module API (MoMA/mem_moco.py, MoMA/criterion_moco_att.py,
 * learning/contrast_trainer.py).  This header is the boundary a binding for that
 * API calls into; every entry point names the reference lines it replaces.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, sizes, scalars, a CUDA stream; no torch types.
 *   - every buffer (inputs, outputs, workspaces, the queue and its bf16 shadow)
 *     is allocated and owned by the caller; the library never allocates or frees
 *     device memory and keeps no pointer past the call.
 *   - all work is enqueued asynchronously on `stream`; no device sync inside.  The
 *     backward entry points (moma_attn_bwd, moma_linear_bwd) fork independent
 *     kernels onto library-owned side streams and join them back into `stream`
 *     with events before returning (parallel branches under graph capture).
 *   - kernels are launched with programmatic stream serialization (PDL): each
 *     waits for its predecessor in `stream` before its first global access, so
 *     the ordering a caller observes is plain stream order.
 *   - return value: MOMA_OK (0) or a negative moma_status; moma_last_error()
 *     returns a thread-local message for the last failure on this thread.
 *   - row-major, fp32 unless a `dtype` argument says otherwise; 16-byte aligned
 *     base pointers and row strides (D % 4 == 0 for fp32, D % 8 == 0 for bf16).
 *   - there is NO CPU fallback: on a machine without a CUDA device every compute
 *     entry point fails with MOMA_ERR_CUDA.
 */
#ifndef MOMA_B200_H_
#define MOMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOMA_ABI_VERSION 3

typedef void *moma_stream_t; /* cudaStream_t / CUstream */

enum moma_status {
    MOMA_OK = 0,
    MOMA_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, n > K ...) */
    MOMA_ERR_ALIGN = -2,       /* pointer / row stride not 16-byte aligned */
    MOMA_ERR_UNSUPPORTED = -3, /* shape outside the kernels' supported set */
    MOMA_ERR_CUDA = -4,        /* CUDA runtime / launch error, or no device */
    MOMA_ERR_WORKSPACE = -5    /* caller-provided workspace too small */
};

enum moma_dtype { MOMA_F32 = 0, MOMA_BF16 = 1 };

int moma_abi_version(void);
const char *moma_last_error(void);
/* 1 when the library was built with the tcgen05/TMA InfoNCE kernel and a device
 * of compute capability 10.x is present. */
int moma_has_tcgen05(void);

/* ------------------------------------------------------------------------- *
 * (c) momentum-encoder EMA
 * replaces learning/contrast_trainer.py:207-211 (momentum_update):
 *     for p1, p2 in zip(model.parameters(), model_ema.parameters()):
 *         p2.data.mul_(m).add_(p1.detach().data, alpha=(1 - m))
 * One launch for the whole parameter list.  Arithmetic is exactly the
 * reference's two fp32 roundings: t = fl(p2*m); p2 = fl(fma(alpha, p1, t)).
 *
 * A "plan" is a caller-owned table describing the tensor list, cut into
 * chunks.  Build it once per (model, model_ema) pair:
 *   1. moma_ema_plan_size   -> number of chunks and table size in bytes
 *   2. moma_ema_plan_fill   -> fill a HOST table (caller copies it to the device)
 *   3. moma_ema_multi       -> every step, with the DEVICE copy of the table
 * ------------------------------------------------------------------------- */
int moma_ema_plan_size(int n_tensors, const int64_t *numels, int64_t *n_chunks, size_t *table_bytes);
int moma_ema_plan_fill(int n_tensors, const void *const *src_ptrs, void *const *dst_ptrs,
                       const int64_t *numels, void *host_table, size_t table_bytes);
int moma_ema_multi(const void *dev_table, int64_t n_chunks, float m, float one_minus_m,
                   moma_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (c) Normalize -- replaces MoMA/criterion_moco_att.py:12-18
 *     F.normalize(x, p=2, dim=1) = x / max(||x||_2, eps)      (and its backward)
 * ------------------------------------------------------------------------- */
int moma_l2norm_fwd(const float *x, float *y, int64_t rows, int64_t D, float eps,
                    moma_stream_t stream);
int moma_l2norm_bwd(const float *x, const float *grad_y, float *grad_x, int64_t rows, int64_t D,
                    float eps, moma_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (c) ring enqueue + pointer -- replaces MoMA/mem_moco.py:14-27
 *     ids = fmod(arange(n) + index, K).long(); queue.index_copy_(0, ids, k)
 *     index = (index + n) % K
 * `K` is the GLOBAL queue length.  With shard_world > 1 the queue is sharded
 * cyclically by row: global row g lives on rank g % shard_world at local slot
 * g / shard_world, `queue_f32`/`queue_bf16` point at this rank's shard
 * [K / shard_world, D], and only owned rows are written (K % shard_world == 0).
 * `queue_bf16` (nullable) is the bf16 shadow read by the tensor-core kernel.
 * `index_dev` (nullable) overrides `index` with a device-resident int64 so the
 * step is CUDA-graph capturable; moma_pointer_advance updates it on the stream.
 * normalize != 0 fuses Normalize (above) into the copy.
 * n > K is rejected (the reference's behaviour is undefined there).
 * ------------------------------------------------------------------------- */
int moma_enqueue(const float *keys, int64_t n, int64_t D, float *queue_f32, void *queue_bf16,
                 int64_t K, int64_t index, const int64_t *index_dev, int shard_rank,
                 int shard_world, int normalize, float eps, moma_stream_t stream);
/* keys[i] is row key_start + i*key_stride of the step's key list: for a rank that only computed the
 * keys it owns (K-sharded queue: every W-th row, see moma_attn_fwd_rows). */
int moma_enqueue_strided(const float *keys, int64_t n, int64_t D, float *queue_f32, void *queue_bf16,
                         int64_t K, int64_t index, const int64_t *index_dev, int shard_rank,
                         int shard_world, int64_t key_start, int64_t key_stride, moma_stream_t stream);
int moma_enqueue_ids(int64_t n, int64_t index, const int64_t *index_dev, int64_t K,
                     int64_t *out_ids, moma_stream_t stream);
int moma_pointer_advance(int64_t *index_dev, int64_t n, int64_t K, moma_stream_t stream);
/* fp32 -> bf16 (round to nearest even); used to (re)build the queue shadow. */
int moma_cast_bf16(const float *src, void *dst_bf16, int64_t numel, moma_stream_t stream);
/* y = x * (*scalar_dev): chain rule of the fused InfoNCE gradient with the upstream scalar (the autograd backward of
 * learning/contrast_trainer.py:197 inside helper/loops_moma.py:345-360), device-resident scalar, no host sync. */
int moma_scale_by_scalar(const float *x, const float *scalar_dev, float *y, int64_t numel, moma_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (a) InfoNCE logits + cross-entropy forward AND backward in one pass
 * replaces MoMA/mem_moco.py:29-49 (_compute_logit), :89 (queue clone),
 * learning/contrast_trainer.py:189-205 (CrossEntropyLoss + top-1 accuracy) and
 * the autograd backward of those ops.  The [B, K+1] logits never reach HBM.
 *
 * Stage 1, moma_nce_partial: for each of `n_splits` contiguous slices of the
 * local queue rows, per query row i
 *     m_i = reference max, l_i = sum_j exp(s_ij - m_i), O_i = sum_j exp(s_ij - m_i) queue_j,
 *     mmax_i = true max_j s_ij,            s_ij = q_i . queue_j / T
 * written to part_m/part_l/part_mmax [n_splits, B] and part_O [n_splits, B, D] (fp32).
 *   dtype MOMA_F32 : q, queue fp32; SIMT FFMA kernel (1e-5 parity mode)
 *   dtype MOMA_BF16: q, queue bf16; tcgen05/TMEM kernel fed by TMA
 * Stage 2, moma_nce_combine: merges `n_parts` partials (splits x shards), adds
 * the positive column q_i.k_i/T and emits
 *     loss_rows[i] = LSE_i - l_i0,   dq_unit[i,:] = (sum_j p_ij c_j - k_i)/T,
 *     pos_is_max[i] = (l_i0 >= max_j l_ij)      (top-1 accuracy by-product)
 * The caller applies the 1/B of CrossEntropyLoss's mean and the upstream grad.
 * q_f32/kpos_f32 are fp32 (in bf16 mode: the bf16-rounded values widened).
 * ------------------------------------------------------------------------- */
int moma_nce_num_splits(int64_t B, int64_t D, int64_t K_local, int dtype);
int moma_nce_partial(const void *q, const void *queue, int64_t B, int64_t D, int64_t K_local,
                     float inv_T, int dtype, int n_splits, float *part_m, float *part_l,
                     float *part_mmax, float *part_O, moma_stream_t stream);
int moma_nce_combine(const float *part_m, const float *part_l, const float *part_mmax,
                     const float *part_O, int n_parts, const float *q_f32, const float *kpos_f32,
                     int64_t B, int64_t D, float inv_T,
                     int round_bf16 /* round q / kpos to bf16 inline (bf16 mode) */,
                     float dq_scale /* folded into dq_unit, e.g. 1/B for the mean */,
                     float *loss_rows, float *dq_unit, int32_t *pos_is_max,
                     float *max_logit /* nullable: max_j l_ij */,
                     float *loss_mean /* nullable pair: mean_i loss_rows[i] ... */,
                     float *acc_pct /* ... and 100 * mean_i pos_is_max[i] (learning/util.py:25-41) */,
                     moma_stream_t stream);
/* Fold `n_parts` partials into ONE partial per row (same (m, l, mmax, O) convention); used by
 * the K-sharded queue before the cross-rank exchange. */
int moma_nce_merge(const float *part_m, const float *part_l, const float *part_mmax,
                   const float *part_O, int n_parts, int64_t B, int64_t D, float *out_m,
                   float *out_l, float *out_mmax, float *out_O, moma_stream_t stream);
/* Packed variants for the cross-rank exchange: one record per row
 *   [ O (D floats) | m | l | mmax | pad ]  = D + 4 floats (rows stay 16-byte aligned),
 * written by moma_nce_merge_packed ([B, D+4]) and consumed by moma_nce_combine_packed
 * ([n_parts, B, D+4], e.g. the all-to-all receive buffer) -- no repacking kernels in between. */
int moma_nce_merge_packed(const float *part_m, const float *part_l, const float *part_mmax,
                          const float *part_O, int n_parts, int64_t B, int64_t D, float *packed,
                          moma_stream_t stream);
int moma_nce_combine_packed(const float *packed, int n_parts, const float *q_f32,
                            const float *kpos_f32, int64_t B, int64_t D, float inv_T, int round_bf16,
                            float dq_scale, float *loss_rows, float *dq_unit, int32_t *pos_is_max,
                            float *max_logit, float *loss_mean, float *acc_pct, moma_stream_t stream);
/* One launch for the whole InfoNCE pass (bf16 operands, D = 64 or 128): the tcgen05 partial kernel, and in its tail the
 * combine -- every CTA waits (bounded) for the other K-splits of its query tile, then combines a slice of the tile's rows;
 * the last CTA of the grid reduces the loss rows to loss_mean / acc_pct.  Same outputs as moma_nce_partial +
 * moma_nce_combine (replaces MoMA/mem_moco.py:29-49,89 + learning/contrast_trainer.py:189-205 and their backward).
 *   workspace : moma_nce_fused_workspace_bytes() bytes, 16-byte aligned (the split partials; L2-resident between the phases)
 *   counters  : >= 256 uint32, ZERO before the first use; the kernel leaves them zero.  Not shared between launches
 *               that may run concurrently.
 * moma_nce_fused_supported: shape check (the grid must fit one CTA per SM: B / 128 * splits <= SM count).
 * moma_nce_fused_packed: same pass, but the tail emits one merged packed record per query row, [B, D + 4] =
 *   (O | m | l | mmax | pad) -- the unit the K-sharded queue exchanges between ranks (then moma_nce_combine_packed). */
int moma_nce_fused_supported(int64_t B, int64_t D, int64_t K_local);
size_t moma_nce_fused_workspace_bytes(int64_t B, int64_t D, int64_t K_local);
int moma_nce_fused(const void *q_bf16, const void *queue_bf16, const float *q_f32, const float *kpos_f32,
                   int64_t B, int64_t D, int64_t K_local, float inv_T, int round_bf16, float dq_scale,
                   void *workspace, size_t workspace_bytes, uint32_t *counters, float *loss_rows,
                   float *dq_unit, int32_t *pos_is_max, float *max_logit, float *loss_mean, float *acc_pct,
                   moma_stream_t stream);
int moma_nce_fused_packed(const void *q_bf16, const void *queue_bf16, int64_t B, int64_t D, int64_t K_local,
                          float inv_T, void *workspace, size_t workspace_bytes, uint32_t *counters,
                          float *packed, moma_stream_t stream);

/* Escape hatch / tests: materialise logits[B, K+1] = cat(q.k, q queue^T) / T
 * exactly as mem_moco.py:29-49 lays them out (row stride K+1). */
int moma_nce_logits(const void *q, const void *kpos, const void *queue, int64_t B, int64_t D,
                    int64_t K, float T, int dtype, float *logits, moma_stream_t stream);
/* positives only: mem_moco.py:51-66 (_compute_logit_qk) -> out[B] */
int moma_nce_logits_qk(const float *q, const float *kpos, int64_t B, int64_t D, float T,
                       float *out, moma_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (b) multi-head attention over the batch axis
 * replaces MoMA/criterion_moco_att.py:153-167 (Attention.forward):
 *     qkv = x W_qkv^T + b_qkv; attn = softmax(q k^T * hd^-0.5); y = (attn v) W_proj^T + b_proj
 * x [N, C]; w_qkv [3C, C]; b_qkv [3C] or NULL; w_proj [C, C]; b_proj [C].
 * Saved for backward (caller-owned): qkv [N, 3C], o [N, C] (merged heads),
 * lse [H, N].  attn_probs (nullable, [H, N, N]) is only for Attention_viz.
 * y_bf16 (nullable, [N, C] bf16): y rounded to bf16, written by the projection's
 * epilogue -- the operand the tcgen05 InfoNCE kernel reads (no separate cast launch).
 * head_dim = C / H must be one of 8, 16, 32, 64, 128.
 * ------------------------------------------------------------------------- */
int moma_attn_fwd(const float *x, const float *w_qkv, const float *b_qkv, const float *w_proj,
                  const float *b_proj, int64_t N, int64_t C, int H, float *y, float *qkv,
                  float *o, float *lse, float *attn_probs, void *y_bf16, moma_stream_t stream);
/* Forward for the query rows q_start + i*q_stride (i < q_count) only; keys / values from all N rows.
 * y, o: [q_count, C]; lse: [H, q_count]; qkv: [N, 3C] scratch.  Forward only (no saved state).
 * x == NULL (then w_qkv / b_qkv are ignored): qkv is an INPUT holding the projections of all N rows,
 * e.g. all-gathered from the ranks that computed them (K-sharded queue: no rank projects all W*B keys). */
int moma_attn_fwd_rows(const float *x, const float *w_qkv, const float *b_qkv, const float *w_proj,
                       const float *b_proj, int64_t N, int64_t C, int H, int64_t q_start,
                       int64_t q_stride, int64_t q_count, float *y, float *qkv, float *o, float *lse,
                       moma_stream_t stream);
size_t moma_attn_bwd_workspace_bytes(int64_t N, int64_t C, int H);
/* Any of the gradient outputs may be NULL (skipped). Gradients are overwritten,
 * not accumulated. */
int moma_attn_bwd(const float *x, const float *w_qkv, const float *w_proj, const float *qkv,
                  const float *o, const float *lse, const float *grad_y, int64_t N, int64_t C,
                  int H, float *grad_x, float *grad_w_qkv, float *grad_b_qkv, float *grad_w_proj,
                  float *grad_b_proj, void *workspace, size_t workspace_bytes,
                  moma_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Projection-head Linear (+ReLU) -- replaces the nn.Linear / nn.ReLU layers of
 * embed_s / embed_t (MoMA/criterion_moco_att.py:254-305) and their autograd.
 *   y[M, N] = act(x[M, K] . w[N, K]^T + b[N]),  act = ReLU when relu != 0;  b nullable.
 * Tensor-core GEMM in 3xTF32 (error-compensated TF32 split, fp32-level accuracy,
 * fp32 accumulation); deterministic split-K through the caller-owned workspace.
 * Workspace: moma_linear_workspace_bytes(M, N, K) bytes, 16-byte aligned, ZEROED
 * once before its first use (its ticket counters are left zero by every call);
 * one workspace per concurrently running call.  NULL workspace = no split-K.
 * Backward: y is the forward output (the ReLU mask; may be NULL when relu == 0);
 * any of grad_x [M, K], grad_w [N, K], grad_b [N] may be NULL; overwritten, not
 * accumulated.
 * ------------------------------------------------------------------------- */
size_t moma_linear_workspace_bytes(int64_t M, int64_t N, int64_t K);
int moma_linear_fwd(const float *x, const float *w, const float *b, int64_t M, int64_t N, int64_t K,
                    int relu, float *y, void *workspace, size_t workspace_bytes,
                    moma_stream_t stream);
int moma_linear_bwd(const float *x, const float *w, const float *y, const float *grad_y, int64_t M,
                    int64_t N, int64_t K, int relu, float *grad_x, float *grad_w, float *grad_b,
                    void *workspace, size_t workspace_bytes, moma_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Peer-memory exchange for the K-sharded queue (one process per GPU, one node).
 * Replaces, on the critical path of a step, the library collectives of
 * learning/contrast_trainer.py:83-88 (_global_gather of the keys) and the query
 * all-gather / partial-record all-to-all of the sharded InfoNCE pass: ONE kernel
 * pushes this rank's rows into every peer's buffer over NVLink (remote stores
 * of 8-byte (payload word, epoch tag) pairs: the flag travels in band, no fence
 * or signal round trip), polls the peers' pairs and writes the payload to `out`.
 * All ranks must issue the same sequence of calls per channel.
 *   peer_bases_dev: device array [world] with the base address of every rank's
 *     copy of one symmetric allocation (same size / layout on all ranks), zeroed
 *     once, holding a control block of moma_peer_ctrl_bytes() bytes at ctrl_off
 *     and, from data_off, 2 regions of region_bytes per channel (4 channels),
 *     region_bytes >= 2 * world * bytes_per_rank.
 *   peer p receives bytes_per_rank bytes from src + p * src_peer_stride_bytes
 *     (stride 0: all-gather; stride = one block: all-to-all); with
 *     cast_f32_to_bf16 the source is fp32 and the pushed rows are bf16.
 *   out: [world, bytes_per_rank], slot s = the rows pushed by rank s.
 * A peer that never arrives traps the kernel after a bounded wait (~30 s; no hang).
 * ------------------------------------------------------------------------- */
size_t moma_peer_ctrl_bytes(void);
int moma_peer_exchange(const void *src, int64_t src_peer_stride_bytes, int64_t bytes_per_rank,
                       int cast_f32_to_bf16, const uint64_t *peer_bases_dev, int64_t ctrl_off,
                       int64_t data_off, int64_t region_bytes, int rank, int world, int channel,
                       void *out, moma_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (f2) optimizer step fused with the momentum-encoder EMA: torch.optim.SGD(momentum, weight_decay; dampening 0, no
 * Nesterov -- train_student_moma.py:389-392) .step() (helper/loops_moma.py:361) followed by momentum_update
 * (learning/contrast_trainer.py:207-211, called at the top of the next iteration, loops_moma.py:309) in ONE multi-tensor
 * pass: 28 B / element instead of 20 + 12 in two passes, one launch instead of ~3 per tensor.  Same rounding sequence as
 * the two library steps (see csrc/sgd_ema.cu).  Plan protocol as for moma_ema_*: size -> fill a HOST table -> copy to
 * the device -> moma_sgd_ema_multi every step (first_step != 0: the momentum buffers are created, buf = d).
 * ------------------------------------------------------------------------- */
int moma_sgd_ema_plan_size(int n_tensors, const int64_t *numels, int64_t *n_chunks, size_t *table_bytes);
int moma_sgd_ema_plan_fill(int n_tensors, void *const *param_ptrs, const void *const *grad_ptrs,
                           void *const *buf_ptrs, void *const *ema_ptrs, const int64_t *numels,
                           void *host_table, size_t table_bytes);
int moma_sgd_ema_multi(const void *dev_table, int64_t n_chunks, float lr, float momentum, float weight_decay,
                       int first_step, float m, float one_minus_m, moma_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (f3) the consumers either side of the MoMA loss on the [B, n_cls] student / teacher logits, one launch:
 * classification CE (train_student_moma.py:294, helper/loops_moma.py:278), Hinton KD
 * DistillKL(T) (distiller_zoo/KD.py:7-17, loops_moma.py:279) and top-1 accuracy (helper/util.py:71-85, :350),
 * as three device scalars out3 = (loss_cls, loss_div, acc_pct) plus the gradients of the two losses with respect to
 * logit_s ([B, n_cls] each), so the autograd backward is a scalar multiply-add.  labels: int64 [B].  n_cls <= 1024.
 * ------------------------------------------------------------------------- */
int moma_cls_kd(const float *logit_s, const float *logit_t, const int64_t *labels, int64_t B, int64_t n_cls,
                float T, float *out3, float *grad_cls, float *grad_div, moma_stream_t stream);

/* K-sharded queue: the exchange of the packed records fused into the merge and combine kernels (protocol and the
 * peer_bases_dev / ctrl_off / data_off / region_bytes / channel arguments of moma_peer_exchange).
 * moma_nce_merge_push: folds the K-splits of all n = world * rows_per_rank query rows (ordered by owner rank) and
 *   stores each merged record straight into the receive region of the rank that owns the query.
 * moma_nce_combine_poll: polls the `world` records of each of this rank's B rows in the local receive region, adds the
 *   positive column and emits loss rows / dq / flags (as moma_nce_combine_packed); advances the channel's epoch.
 * Every rank calls both, in this order, on the same channel, once per step; D % 4 == 0, D <= 512. */
int moma_nce_merge_push(const float *part_m, const float *part_l, const float *part_mmax, const float *part_O,
                        int n_parts, int64_t rows_per_rank, int64_t D, const uint64_t *peer_bases_dev,
                        int64_t ctrl_off, int64_t data_off, int64_t region_bytes, int rank, int world, int channel,
                        moma_stream_t stream);
int moma_nce_combine_poll(const float *q_f32, const float *kpos_f32, int64_t B, int64_t D, float inv_T,
                          int round_bf16, float dq_scale, const uint64_t *peer_bases_dev, int64_t ctrl_off,
                          int64_t data_off, int64_t region_bytes, int rank, int world, int channel,
                          float *loss_rows, float *dq_unit, int32_t *pos_is_max, float *max_logit,
                          float *loss_mean, float *acc_pct, moma_stream_t stream);

/* ------------------------------------------------------------------------- *
 * debug / test hooks (not used by the product path)
 * moma_debug_nce_tc: the tcgen05 partial kernel with an optional dump of the raw
 *   score tile S = Q . Tile^T of the first queue tile of split 0:
 *   dbg_S [B, BN] (BN = 128, or 64 when D == 256), nullable.
 * moma_debug_tc_error: synchronising read of the device-side protocol-error code
 *   (a bounded mbarrier wait that timed out records which one and traps); 0 = none.
 * ------------------------------------------------------------------------- */
int moma_debug_nce_tc(const void *q, const void *queue, int64_t B, int64_t D, int64_t K_local,
                      float inv_T, int n_splits, float *part_m, float *part_l, float *part_mmax,
                      float *part_O, float *dbg_S, moma_stream_t stream);
int moma_debug_tc_error(void);
/* number of kernels this library has launched in this process (optionally reset to 0) */
long long moma_debug_launch_count(int reset);
/* ALGORITHMIC floating-point operations issued since the last reset (for bench.py's rooflines):
 * kind 0 = Linear GEMMs of moma_linear_* / the attention projections (2*M*N*K each, counted once although the
 * 3xTF32 kernel runs three tensor-core passes), kind 1 = attention core (forward 4*Nq*N*C, backward 8*N*N*C). */
double moma_debug_flops(int kind, int reset);
/* Turn programmatic dependent launch on / off for subsequent launches (returns the previous setting).  A kernel
 * launched with the attribute starts while its predecessor drains and then waits, so profiler "durations" include
 * that wait: bench.py captures a second graph with PDL off to measure per-kernel shares. */
int moma_debug_set_pdl(int enable);
/* Launch-overhead probe: an empty kernel of `ctas` x `threads` with `smem_bytes` of dynamic shared memory, optionally
 * allocating / releasing all TMEM columns (the launch shape of the InfoNCE kernel, without its work). */
int moma_debug_probe_launch(int ctas, int threads, int smem_bytes, int use_tmem, int pdl, moma_stream_t stream);

/* The tcgen05 3xTF32 GEMM of csrc/gemm_tc.cu on its own (tests): C[M,N] = act((A o (mask > 0)) . B + bias) with
 * A stored [M,K] (a_mn = 0) or [K,M] (a_mn = 1) and B stored [N,K] (b_mn = 0) or [K,N] (b_mn = 1); lda / ldb are the
 * row strides of the stored matrices; a_mn = 1 with b_mn = 0 is not built.  mask (nullable) has A's stored shape and
 * stride.  workspace from
 * moma_debug_gemm_tc_workspace_bytes (zero-filled before first use; NULL = no split-K).  MOMA_ERR_UNSUPPORTED for
 * shapes the kernel does not take (M < 128, unaligned strides). */
int moma_debug_gemm_tc(const float *A, int64_t lda, int a_mn, const float *mask, const float *B, int64_t ldb, int b_mn,
                       const float *bias, float *C, int64_t ldc, int64_t M, int64_t N, int64_t K, int relu,
                       void *workspace, size_t workspace_bytes, moma_stream_t stream);
size_t moma_debug_gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K);
int moma_debug_gemm_tc_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MOMA_B200_H_ */
