// InfoNCE logits + cross-entropy, FP32 parity path (SIMT FFMA) + combine + materialise.
//
// Replaces MoMA/mem_moco.py:29-49 (_compute_logit), the CrossEntropyLoss/top-1 at
// learning/contrast_trainer.py:189-205 and their autograd backward.
//
// Algebra: labels are all zero, so with s_ij = q_i.queue_j / T
//   loss_i = LSE_j(l_ij) - l_i0,   dloss_i/dq_i = (sum_j p_ij c_j - k_i) / T,
// i.e. one flash-style pass over the queue yields both the LSE and P.queue.
// moma_nce_partial produces per-(split) partials (m, l, O); moma_nce_combine
// merges splits and shards, adds the positive column and emits loss rows + dq.
//
// The FP32 kernel here is the 1e-5-parity mode (no tensor cores: a single TF32
// pass fails 1e-5 on the gradient).  The bf16 tensor-core
// kernel lives in nce_tc.cu and shares the combine below.
#include <math_constants.h>
#include "common.cuh"

namespace moma {

int nce_tc_partial(const void* q, const void* queue, int64_t B, int64_t D, int64_t K_local,
                   float inv_T, int n_splits, float* part_m, float* part_l, float* part_mmax,
                   float* part_O, cudaStream_t stream);        // nce_tc.cu
int nce_tc_num_splits(int64_t B, int64_t D, int64_t K_local);  // nce_tc.cu
bool nce_tc_supported(int64_t B, int64_t D, int64_t K_local);  // nce_tc.cu
bool nce_fused_supported(int64_t B, int64_t D, int64_t K_local);
size_t nce_fused_workspace_floats(int64_t B, int64_t D, int64_t K_local);
int nce_fused_launch(const void* q_bf16, const void* queue_bf16, int64_t B, int64_t D, int64_t K_local, float inv_T,
                     float* workspace, const tc::NceFuse& fuse, cudaStream_t stream);

// ------------------------------------------------------------------ fp32 partial
// CTA = 256 threads = 8 warps; BM = 32 query rows (4 per warp), BN queue rows per
// tile.  Q tile and queue tile live in shared memory with a +4 float row pad so
// 128-bit row reads by consecutive lanes are bank-conflict free.
constexpr int kBM = 32;
constexpr int kSimtThreads = 256;

template <int BN>   // 64 (D <= 256) or 32 (D <= 512)
__global__ void __launch_bounds__(kSimtThreads)
nce_partial_f32_kernel(const float* __restrict__ q, const float* __restrict__ queue, int B, int D,
                       int64_t K, float inv_T, int n_splits, float* __restrict__ part_m,
                       float* __restrict__ part_l, float* __restrict__ part_mmax,
                       float* __restrict__ part_O) {
    extern __shared__ __align__(16) float smem[];
    const int ld = D + 4;
    float* qs = smem;                       // [kBM][ld]
    float* ks = qs + kBM * ld;              // [BN][ld]
    float* ps = ks + BN * ld;               // [kBM][BN + 4]
    constexpr int ldp = BN + 4;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * kBM;
    const int split = blockIdx.y;
    // split -> contiguous range of BN-row tiles
    const int64_t n_tiles = (K + BN - 1) / BN;
    const int64_t t_begin = n_tiles * split / n_splits;
    const int64_t t_end = n_tiles * (split + 1) / n_splits;

    const int nvec = D >> 2;
    for (int i = tid; i < kBM * nvec; i += kSimtThreads) {
        const int r = i / nvec, v = i - r * nvec;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + r < B) a = reinterpret_cast<const float4*>(q + (int64_t)(m0 + r) * D)[v];
        *reinterpret_cast<float4*>(qs + r * ld + 4 * v) = a;
    }

    // running stats for this warp's 4 rows (replicated across lanes)
    float m_run[4], l_run[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { m_run[r] = -CUDART_INF_F; l_run[r] = 0.f; }
    // O accumulators: 4 rows x (D/32) columns, column = lane + 32*c   (D <= 512 -> c < 16)
    constexpr int kMaxC = 16;
    float acc[4][kMaxC];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < kMaxC; ++c) acc[r][c] = 0.f;
    const int ncol = (D + 31) >> 5;

    for (int64_t t = t_begin; t < t_end; ++t) {
        const int64_t j0 = t * BN;
        __syncthreads();                                   // previous tile fully consumed
        for (int i = tid; i < BN * nvec; i += kSimtThreads) {
            const int r = i / nvec, v = i - r * nvec;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j0 + r < K) a = ld_stream(reinterpret_cast<const float4*>(queue + (j0 + r) * D) + v);
            *reinterpret_cast<float4*>(ks + r * ld + 4 * v) = a;
        }
        __syncthreads();

        // S[4 rows][BN/32 cols per lane]: column j = lane + 32*cc
        constexpr int CC = BN / 32;
        float s[4][CC];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) s[r][cc] = 0.f;
        for (int v = 0; v < nvec; ++v) {
            float4 kk[CC];
#pragma unroll
            for (int cc = 0; cc < CC; ++cc)
                kk[cc] = *reinterpret_cast<const float4*>(ks + (lane + 32 * cc) * ld + 4 * v);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 qq = *reinterpret_cast<const float4*>(qs + (warp * 4 + r) * ld + 4 * v);
#pragma unroll
                for (int cc = 0; cc < CC; ++cc)
                    s[r][cc] += qq.x * kk[cc].x + qq.y * kk[cc].y + qq.z * kk[cc].z + qq.w * kk[cc].w;
            }
        }
        // online softmax per row (warp-shuffle reductions)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float mx = -CUDART_INF_F;
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
                s[r][cc] = (j0 + lane + 32 * cc < K) ? s[r][cc] * inv_T : -CUDART_INF_F;
                mx = fmaxf(mx, s[r][cc]);
            }
            mx = warp_max(mx);
            const float m_new = fmaxf(m_run[r], mx);
            const float corr = (m_run[r] == -CUDART_INF_F) ? 0.f : expf(m_run[r] - m_new);
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
                const float p = (s[r][cc] == -CUDART_INF_F) ? 0.f : expf(s[r][cc] - m_new);
                ps[(warp * 4 + r) * ldp + lane + 32 * cc] = p;
                sum += p;
            }
            sum = warp_sum(sum);
            l_run[r] = l_run[r] * corr + sum;
            m_run[r] = m_new;
#pragma unroll
            for (int c = 0; c < kMaxC; ++c)
                if (c < ncol) acc[r][c] *= corr;
        }
        __syncwarp();
        // O += P . tile
        for (int j = 0; j < BN; j += 4) {
            float4 p4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
                p4[r] = *reinterpret_cast<const float4*>(ps + (warp * 4 + r) * ldp + j);
#pragma unroll
            for (int c = 0; c < kMaxC; ++c) {
                if (c < ncol) {
                    const int col = lane + 32 * c;
                    const bool ok = col < D;
                    const float v0 = ok ? ks[(j + 0) * ld + col] : 0.f;
                    const float v1 = ok ? ks[(j + 1) * ld + col] : 0.f;
                    const float v2 = ok ? ks[(j + 2) * ld + col] : 0.f;
                    const float v3 = ok ? ks[(j + 3) * ld + col] : 0.f;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        acc[r][c] += p4[r].x * v0 + p4[r].y * v1 + p4[r].z * v2 + p4[r].w * v3;
                }
            }
        }
    }

    // write partials
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int row = m0 + warp * 4 + r;
        if (row >= B) continue;
        const int64_t o = (int64_t)split * B + row;
        if (lane == 0) { part_m[o] = m_run[r]; part_l[o] = l_run[r]; part_mmax[o] = m_run[r]; }
#pragma unroll
        for (int c = 0; c < kMaxC; ++c) {
            const int col = lane + 32 * c;
            if (c < ncol && col < D) part_O[o * D + col] = acc[r][c];
        }
    }
}

// ------------------------------------------------------------------ combine / merge
// One CTA of 128 threads (4 warps) per query row.  Scalar reductions over the partials (max, weights,
// l) are spread over the threads; the O accumulation splits the PARTS over the 4 warps, each lane owning
// a float4 of columns, with 8 independent 128-bit loads in flight per lane (the kernel is latency-bound:
// n_parts * D * 4 bytes per row come from L2), then one cross-warp reduction through shared memory.
constexpr int kCombThreads = 128;
constexpr int kCombChunk = 512;       // parts staged per pass
constexpr int kCombMaxD = 512;        // columns per pass (4 float4 per lane)

__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    v = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    return v;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    v = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
    return v;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <bool kShared>
__device__ __forceinline__ float4 ld_part(const float* p) {       // partial O: L2-resident global memory, or staged in shared memory
    if (kShared) return *reinterpret_cast<const float4*>(p);
    return __ldg(reinterpret_cast<const float4*>(p));
}

// kFinal: add the positive column and emit loss / dq / flags; otherwise emit one merged partial.
// kPeer (K-sharded queue): the exchange between the ranks happens INSIDE the two kernels, in the protocol of
// csrc/peer.cu -- no exchange launch between them:
//   merge  (!kFinal): the merged record [O | m | l | mmax | pad] of global query row r is stored straight into the
//                     receive region of the rank that owns the query (r / rows_per_rank), as (word, epoch tag) cells;
//   combine (kFinal): polls the `world` records of its row in the local receive region into shared memory and folds
//                     them; the last CTA advances the channel's epoch.
template <bool kFinal, bool kPeer>
__global__ void __launch_bounds__(kCombThreads)
nce_reduce_kernel(const float* __restrict__ part_m_, const float* __restrict__ part_l_,
                  const float* __restrict__ part_mmax_, const float* __restrict__ part_O_, int n_parts,
                  const float* __restrict__ q, const float* __restrict__ kpos, int B, int D, float inv_T,
                  int round_bf16, float dq_scale,
                  int64_t st_ss_ /* stats stride between parts */, int64_t st_rs_ /* ... between rows */,
                  int64_t o_ss_ /* O stride between parts */, int64_t o_rs_ /* ... between rows */,
                  int64_t out_rs /* merged-partial output: row stride of the stats (1 = separate arrays) */,
                  int64_t out_o_rs /* row stride of the O output */,
                  float* __restrict__ out_a /* loss_rows | out_m */, float* __restrict__ out_O /* dq | out_O */,
                  int32_t* __restrict__ pos_is_max, float* __restrict__ out_b /* max_logit | out_l */,
                  float* __restrict__ out_c /* - | out_mmax */, const PeerLink link) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[4];
    __shared__ float s_w[kCombChunk];
    __shared__ __align__(16) float s_o[4][kCombMaxD];
    extern __shared__ __align__(16) float s_in[];                  // kPeer && kFinal: [world][D + 4] received records
    const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* kr = kFinal ? kpos + (int64_t)row * D : nullptr;
    const float *part_m = part_m_, *part_l = part_l_, *part_mmax = part_mmax_, *part_O = part_O_;
    int64_t st_ss = st_ss_, st_rs = st_rs_, o_ss = o_ss_, o_rs = o_rs_;

    PeerCtrl* ctrl = nullptr;
    unsigned long long epoch = 0ull;
    long long region = 0;
    if (kPeer) {
        ctrl = reinterpret_cast<PeerCtrl*>(link.bases[link.rank] + link.ctrl_off);
        epoch = *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch[link.channel]) + 1ull;
        region = link.data_off + (2ll * link.channel + (long long)(epoch & 1ull)) * link.region_bytes;
    }
    const uint32_t tag = (uint32_t)epoch;
    const int cells = (D + 4) / 2;                                 // 16-byte (w0, tag, w1, tag) cells per record
    if (kPeer && kFinal) {
        // poll this row's record from every rank: slot s of the local region holds the rows_per_rank records rank s pushed
        const uint4* in = reinterpret_cast<const uint4*>(link.bases[link.rank] + region);
        const long long t0 = clock64();
        // One thread per source rank sleeps on the LAST cell of that rank's record; the other threads park at the barrier.
        // (All 128 threads of all B CTAs spinning on their own cells was measured at 8 GPUs: lower best-case step, but the
        // polling traffic slowed the step's other branches -- median 0.315 ms against 0.287 with the separate exchange.)
        if (tid < link.world) {
            const uint4* cell = in + ((long long)tid * link.rows_per_rank + row) * cells + (cells - 1);
            unsigned ns = 32;
            uint4 v = ld_volatile16(cell);
            while (v.y != tag || v.w != tag) {
                __nanosleep(ns);
                if (ns < 512) ns *= 2;
                if (clock64() - t0 > kPeerTimeoutClk) __trap();
                v = ld_volatile16(cell);
            }
        }
        __syncthreads();
        for (int i = tid; i < link.world * cells; i += kCombThreads) {
            const int s = i / cells, c = i - s * cells;
            const uint4* cell = in + ((long long)s * link.rows_per_rank + row) * cells + c;
            uint4 v = ld_volatile16(cell);
            while (v.y != tag || v.w != tag) {
                if (clock64() - t0 > kPeerTimeoutClk) __trap();
                v = ld_volatile16(cell);
            }
            s_in[s * (D + 4) + 2 * c] = __uint_as_float(v.x);
            s_in[s * (D + 4) + 2 * c + 1] = __uint_as_float(v.z);
        }
        __syncthreads();
        part_O = s_in; part_m = s_in + D; part_l = s_in + D + 1; part_mmax = s_in + D + 2;
        st_ss = o_ss = D + 4; st_rs = o_rs = 0;
    }

    float pos = 0.f;
    if (kFinal) {
        const float* qr = q + (int64_t)row * D;
        float dot = 0.f;
        for (int d = tid; d < D; d += kCombThreads) {
            const float qv = round_bf16 ? bf16_round(qr[d]) : qr[d];
            const float kv = round_bf16 ? bf16_round(kr[d]) : kr[d];
            dot += qv * kv;
        }
        pos = block_sum(dot, red) * inv_T;
    }
    float mref = -CUDART_INF_F, mtrue = -CUDART_INF_F;
    for (int s = tid; s < n_parts; s += kCombThreads) {
        mref = fmaxf(mref, part_m[(int64_t)s * st_ss + row * st_rs]);
        mtrue = fmaxf(mtrue, part_mmax[(int64_t)s * st_ss + row * st_rs]);
    }
    mref = block_max(mref, red);
    mtrue = block_max(mtrue, red);
    const float mstar = kFinal ? fmaxf(mref, pos) : mref;
    const float wpos = kFinal ? expf(pos - mstar) : 0.f;

    float l_part = 0.f, l_tot = 0.f;
    const int64_t sstride = o_ss;
    for (int d0 = 0; d0 < D; d0 += kCombMaxD) {
        float4 o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < n_parts; s0 += kCombChunk) {
            const int cnt = min(kCombChunk, n_parts - s0);
            __syncthreads();
            for (int s = tid; s < cnt; s += kCombThreads) {
                const int64_t base = (int64_t)(s0 + s) * st_ss + row * st_rs;
                const float ms = part_m[base];
                const float w = (ms == -CUDART_INF_F) ? 0.f : expf(ms - mstar);
                s_w[s] = w;
                if (d0 == 0) l_part += w * part_l[base];
            }
            __syncthreads();
            const float* Op = part_O + (int64_t)s0 * o_ss + row * o_rs + d0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int d = 4 * lane + 128 * c;              // column of this lane's float4
                if (d0 + d < D) {
                    float4 acc = o[c];
                    int s = warp;
                    for (; s + 28 < cnt; s += 32) {            // 8 parts of this warp per round
                        float4 v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            v[u] = ld_part<kPeer && kFinal>(Op + (int64_t)(s + 4 * u) * sstride + d);
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float w = s_w[s + 4 * u];
                            acc.x = fmaf(w, v[u].x, acc.x); acc.y = fmaf(w, v[u].y, acc.y);
                            acc.z = fmaf(w, v[u].z, acc.z); acc.w = fmaf(w, v[u].w, acc.w);
                        }
                    }
                    for (; s < cnt; s += 4) {
                        const float4 v = ld_part<kPeer && kFinal>(Op + (int64_t)s * sstride + d);
                        const float w = s_w[s];
                        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
                        acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                    }
                    o[c] = acc;
                }
            }
        }
        if (d0 == 0) l_tot = block_sum(l_part, red) + wpos;
        // cross-warp reduction of the O partial sums
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int d = 4 * lane + 128 * c;
            if (d0 + d < D) *reinterpret_cast<float4*>(&s_o[warp][d]) = o[c];
        }
        __syncthreads();
        for (int d = tid; d < kCombMaxD && d0 + d < D; d += kCombThreads) {
            const float ov = (s_o[0][d] + s_o[1][d]) + (s_o[2][d] + s_o[3][d]);
            if (kFinal) {
                const float kv = round_bf16 ? bf16_round(kr[d0 + d]) : kr[d0 + d];
                out_O[(int64_t)row * out_o_rs + d0 + d] = ((ov + wpos * kv) / l_tot - kv) * inv_T * dq_scale;
            } else if (kPeer) {
                s_o[0][d] = ov;                                    // (this thread is the only reader of column d)
            } else {
                out_O[(int64_t)row * out_o_rs + d0 + d] = ov;
            }
        }
    }
    if (kPeer && !kFinal) {
        // push the record (D <= kCombMaxD: one pass) to the owner of the query, as tagged cells
        __syncthreads();
        const int owner = row / link.rows_per_rank, local = row - owner * link.rows_per_rank;
        uint4* dst = reinterpret_cast<uint4*>(link.bases[owner] + region) + ((long long)link.rank * link.rows_per_rank + local) * cells;
        for (int c = tid; c < cells; c += kCombThreads) {
            float w0, w1;
            if (2 * c < D) { w0 = s_o[0][2 * c]; w1 = s_o[0][2 * c + 1]; }
            else if (2 * c == D) { w0 = mref; w1 = l_tot; }
            else { w0 = mtrue; w1 = 0.f; }
            dst[c] = make_uint4(__float_as_uint(w0), tag, __float_as_uint(w1), tag);
        }
        return;
    }
    if (tid == 0) {
        if (kFinal) {
            out_a[row] = logf(l_tot) + mstar - pos;
            pos_is_max[row] = (pos >= mtrue) ? 1 : 0;
            if (out_b) out_b[row] = fmaxf(pos, mtrue);
        } else {
            out_a[row * out_rs] = mref; out_b[row * out_rs] = l_tot; out_c[row * out_rs] = mtrue;
        }
    }
    if (kPeer && kFinal) {
        __syncthreads();
        if (tid == 0 && atomicAdd(&ctrl->ticket_done[link.channel], 1u) == gridDim.x - 1u) {   // every CTA has read the epoch
            ctrl->ticket_done[link.channel] = 0u;
            *reinterpret_cast<volatile unsigned long long*>(&ctrl->epoch[link.channel]) = epoch;
        }
    }
}

// loss = mean(rows), acc = 100 * mean(pos_is_max): one block, deterministic order
__global__ void __launch_bounds__(256)
nce_finalize_kernel(const float* __restrict__ rows, const int32_t* __restrict__ pim, int B,
                    float* __restrict__ loss_mean, float* __restrict__ acc_pct) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float r1[8], r2[8];
    float s = 0.f, c = 0.f;
    for (int i = threadIdx.x; i < B; i += 256) { s += rows[i]; c += (float)pim[i]; }
    s = warp_sum(s); c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s; r2[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ts = 0.f, tc = 0.f;
        for (int i = 0; i < 8; ++i) { ts += r1[i]; tc += r2[i]; }
        *loss_mean = ts / (float)B;
        *acc_pct = tc * (100.0f / (float)B);
    }
}

// ------------------------------------------------------------------ materialise
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// logits[i, 1 + j] = q_i . queue_j / T ; 64x64 tile, 16-wide k-step, 4x4 per thread
template <typename T>
__global__ void __launch_bounds__(256)
nce_logits_kernel(const T* __restrict__ q, const T* __restrict__ queue, int B, int D, int64_t K,
                  float Tm, float* __restrict__ logits) {
    __shared__ float qs[16][64 + 1];
    __shared__ float ks[16][64 + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64;
    const int64_t n0 = (int64_t)blockIdx.x * 64;
    float acc[4][4] = {};
    for (int d0 = 0; d0 < D; d0 += 16) {
        for (int i = tid; i < 64 * 16; i += 256) {
            const int r = i >> 4, d = i & 15;
            qs[d][r] = (m0 + r < B && d0 + d < D) ? to_f<T>(q[(int64_t)(m0 + r) * D + d0 + d]) : 0.f;
            ks[d][r] = (n0 + r < K && d0 + d < D) ? to_f<T>(queue[(n0 + r) * D + d0 + d]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = qs[d][ty + 16 * i]; b[i] = ks[d][tx + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = m0 + ty + 16 * i;
        if (r >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = n0 + tx + 16 * j;
            if (c < K) logits[(int64_t)r * (K + 1) + 1 + c] = acc[i][j] / Tm;
        }
    }
}

// column 0 (or the whole output of _compute_logit_qk): one warp per row
template <typename T>
__global__ void __launch_bounds__(128)
nce_pos_kernel(const T* __restrict__ q, const T* __restrict__ kpos, int B, int D, float Tm,
               float* __restrict__ out, int64_t out_stride) {
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B) return;
    float dot = 0.f;
    for (int d = lane; d < D; d += 32)
        dot += to_f<T>(q[(int64_t)row * D + d]) * to_f<T>(kpos[(int64_t)row * D + d]);
    dot = warp_sum(dot);
    if (lane == 0) out[(int64_t)row * out_stride] = dot / Tm;
}

static int partial_f32(const float* q, const float* queue, int64_t B, int64_t D, int64_t K,
                       float inv_T, int n_splits, float* pm, float* pl, float* pmm, float* pO,
                       cudaStream_t st) {
    MOMA_REQUIRE(D % 4 == 0 && D <= 512, MOMA_ERR_UNSUPPORTED,
                 "nce_partial(f32): D=%lld unsupported (need D %% 4 == 0, D <= 512); use moma_nce_logits",
                 (long long)D);
    const dim3 grid((unsigned)((B + kBM - 1) / kBM), (unsigned)n_splits);
    const int ld = (int)D + 4;
    if (D <= 256) {
        constexpr int BN = 64;
        const size_t smem = (size_t)(kBM * ld + BN * ld + kBM * (BN + 4)) * sizeof(float);
        ensure_dyn_smem(reinterpret_cast<const void*>(nce_partial_f32_kernel<BN>), 160 * 1024);
        nce_partial_f32_kernel<BN><<<grid, kSimtThreads, smem, st>>>(q, queue, (int)B, (int)D, K, inv_T, n_splits, pm, pl, pmm, pO);
    } else {
        constexpr int BN = 32;
        const size_t smem = (size_t)(kBM * ld + BN * ld + kBM * (BN + 4)) * sizeof(float);
        ensure_dyn_smem(reinterpret_cast<const void*>(nce_partial_f32_kernel<BN>), 160 * 1024);
        nce_partial_f32_kernel<BN><<<grid, kSimtThreads, smem, st>>>(q, queue, (int)B, (int)D, K, inv_T, n_splits, pm, pl, pmm, pO);
    }
    MOMA_CUDA_LAUNCH_CHECK("nce_partial(f32)");
    note_launches(1);
    return MOMA_OK;
}

}  // namespace moma

using namespace moma;

extern "C" __attribute__((visibility("default"))) int moma_nce_num_splits(int64_t B, int64_t D, int64_t K_local, int dtype) {
    if (B <= 0 || D <= 0 || K_local <= 0) return 1;
    if (dtype == MOMA_BF16 && nce_tc_supported(B, D, K_local)) return nce_tc_num_splits(B, D, K_local);
    // fp32 SIMT: aim for ~2 CTAs per SM, at least 4 tiles of work per split
    const int64_t mt = (B + kBM - 1) / kBM;
    const int64_t bn = D <= 256 ? 64 : 32;
    const int64_t tiles = (K_local + bn - 1) / bn;
    int64_t s = (2 * (int64_t)sm_count() + mt - 1) / mt;
    if (s > tiles / 4) s = tiles / 4;
    if (s < 1) s = 1;
    if (s > 1024) s = 1024;
    return (int)s;
}

extern "C" __attribute__((visibility("default"))) int moma_nce_partial(const void* q, const void* queue, int64_t B, int64_t D,
                                int64_t K_local, float inv_T, int dtype, int n_splits,
                                float* part_m, float* part_l, float* part_mmax, float* part_O,
                                moma_stream_t stream) {
    MOMA_REQUIRE(B > 0 && D > 0 && K_local > 0 && n_splits > 0, MOMA_ERR_INVALID,
                 "nce_partial: bad shape B=%lld D=%lld K=%lld splits=%d", (long long)B, (long long)D,
                 (long long)K_local, n_splits);
    MOMA_REQUIRE(q && queue && part_m && part_l && part_mmax && part_O, MOMA_ERR_INVALID, "nce_partial: null pointer");
    MOMA_REQUIRE(aligned16(q) && aligned16(queue) && aligned16(part_O), MOMA_ERR_ALIGN, "nce_partial: unaligned pointer");
    MOMA_REQUIRE(B < (1ll << 30) && K_local < (1ll << 40), MOMA_ERR_UNSUPPORTED, "nce_partial: shape too large");
    if (dtype == MOMA_F32)
        return partial_f32(static_cast<const float*>(q), static_cast<const float*>(queue), B, D, K_local,
                           inv_T, n_splits, part_m, part_l, part_mmax, part_O, as_stream(stream));
    MOMA_REQUIRE(dtype == MOMA_BF16, MOMA_ERR_INVALID, "nce_partial: unknown dtype %d", dtype);
    MOMA_REQUIRE(nce_tc_supported(B, D, K_local), MOMA_ERR_UNSUPPORTED,
                 "nce_partial(bf16): D=%lld unsupported by the tcgen05 kernel (need D in {64,128,256})",
                 (long long)D);
    return nce_tc_partial(q, queue, B, D, K_local, inv_T, n_splits, part_m, part_l, part_mmax, part_O,
                          as_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int moma_nce_combine(const float* part_m, const float* part_l, const float* part_mmax,
                                const float* part_O, int n_parts, const float* q_f32,
                                const float* kpos_f32, int64_t B, int64_t D, float inv_T,
                                int round_bf16, float dq_scale, float* loss_rows, float* dq_unit,
                                int32_t* pos_is_max, float* max_logit, float* loss_mean, float* acc_pct,
                                moma_stream_t stream) {
    MOMA_REQUIRE(B > 0 && D > 0 && n_parts > 0, MOMA_ERR_INVALID, "nce_combine: bad shape");
    MOMA_REQUIRE(part_m && part_l && part_mmax && part_O && q_f32 && kpos_f32 && loss_rows && dq_unit && pos_is_max,
                 MOMA_ERR_INVALID, "nce_combine: null pointer");
    MOMA_REQUIRE(D % 4 == 0 && aligned16(part_O), MOMA_ERR_ALIGN, "nce_combine: D %% 4 != 0 or part_O unaligned");
    launch_pdl(nce_reduce_kernel<true, false>, dim3((unsigned)B), dim3(kCombThreads), 0, as_stream(stream),
        part_m, part_l, part_mmax, part_O, n_parts, q_f32, kpos_f32, (int)B, (int)D, inv_T, round_bf16, dq_scale,
        B, 1, B * D, D, 1, D, loss_rows, dq_unit, pos_is_max, max_logit, nullptr, PeerLink{});
    MOMA_CUDA_LAUNCH_CHECK("nce_combine");
    note_launches(1);
    if (loss_mean && acc_pct) {
        launch_pdl(nce_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), loss_rows, pos_is_max, (int)B, loss_mean, acc_pct);
        MOMA_CUDA_LAUNCH_CHECK("nce_finalize");
        note_launches(1);
    }
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_nce_merge(
    const float* part_m, const float* part_l, const float* part_mmax, const float* part_O, int n_parts,
    int64_t B, int64_t D, float* out_m, float* out_l, float* out_mmax, float* out_O, moma_stream_t stream) {
    MOMA_REQUIRE(B > 0 && D > 0 && n_parts > 0, MOMA_ERR_INVALID, "nce_merge: bad shape");
    MOMA_REQUIRE(part_m && part_l && part_mmax && part_O && out_m && out_l && out_mmax && out_O,
                 MOMA_ERR_INVALID, "nce_merge: null pointer");
    MOMA_REQUIRE(D % 4 == 0 && aligned16(part_O), MOMA_ERR_ALIGN, "nce_merge: D %% 4 != 0 or part_O unaligned");
    launch_pdl(nce_reduce_kernel<false, false>, dim3((unsigned)B), dim3(kCombThreads), 0, as_stream(stream),
        part_m, part_l, part_mmax, part_O, n_parts, nullptr, nullptr, (int)B, (int)D, 1.f, 0, 1.f,
        B, 1, B * D, D, 1, D, out_m, out_O, nullptr, out_l, out_mmax, PeerLink{});
    MOMA_CUDA_LAUNCH_CHECK("nce_merge");
    note_launches(1);
    return MOMA_OK;
}

// Packed partial record per row: [O (D floats) | m | l | mmax | pad] = D + 4 floats (16-byte aligned rows),
// the unit that travels between ranks for the K-sharded queue.
extern "C" __attribute__((visibility("default"))) int moma_nce_merge_packed(
    const float* part_m, const float* part_l, const float* part_mmax, const float* part_O, int n_parts,
    int64_t B, int64_t D, float* packed, moma_stream_t stream) {
    MOMA_REQUIRE(B > 0 && D > 0 && n_parts > 0, MOMA_ERR_INVALID, "nce_merge_packed: bad shape");
    MOMA_REQUIRE(part_m && part_l && part_mmax && part_O && packed, MOMA_ERR_INVALID, "nce_merge_packed: null pointer");
    MOMA_REQUIRE(D % 4 == 0 && aligned16(part_O) && aligned16(packed), MOMA_ERR_ALIGN, "nce_merge_packed: alignment");
    const int64_t P = D + 4;
    launch_pdl(nce_reduce_kernel<false, false>, dim3((unsigned)B), dim3(kCombThreads), 0, as_stream(stream),
        part_m, part_l, part_mmax, part_O, n_parts, nullptr, nullptr, (int)B, (int)D, 1.f, 0, 1.f,
        B, 1, B * D, D, P, P, packed + D, packed, nullptr, packed + D + 1, packed + D + 2, PeerLink{});
    MOMA_CUDA_LAUNCH_CHECK("nce_merge_packed");
    note_launches(1);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_nce_combine_packed(
    const float* packed, int n_parts, const float* q_f32, const float* kpos_f32, int64_t B, int64_t D,
    float inv_T, int round_bf16, float dq_scale, float* loss_rows, float* dq_unit, int32_t* pos_is_max,
    float* max_logit, float* loss_mean, float* acc_pct, moma_stream_t stream) {
    MOMA_REQUIRE(B > 0 && D > 0 && n_parts > 0, MOMA_ERR_INVALID, "nce_combine_packed: bad shape");
    MOMA_REQUIRE(packed && q_f32 && kpos_f32 && loss_rows && dq_unit && pos_is_max, MOMA_ERR_INVALID,
                 "nce_combine_packed: null pointer");
    MOMA_REQUIRE(D % 4 == 0 && aligned16(packed), MOMA_ERR_ALIGN, "nce_combine_packed: alignment");
    const int64_t P = D + 4;
    launch_pdl(nce_reduce_kernel<true, false>, dim3((unsigned)B), dim3(kCombThreads), 0, as_stream(stream),
        packed + D, packed + D + 1, packed + D + 2, packed, n_parts, q_f32, kpos_f32, (int)B, (int)D, inv_T,
        round_bf16, dq_scale, B * P, P, B * P, P, 1, D, loss_rows, dq_unit, pos_is_max, max_logit, nullptr, PeerLink{});
    MOMA_CUDA_LAUNCH_CHECK("nce_combine_packed");
    note_launches(1);
    if (loss_mean && acc_pct) {
        launch_pdl(nce_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), loss_rows, pos_is_max, (int)B, loss_mean, acc_pct);
        MOMA_CUDA_LAUNCH_CHECK("nce_finalize");
        note_launches(1);
    }
    return MOMA_OK;
}

// ---- K-sharded queue: merge + push / poll + combine (the exchange of the packed records runs inside the two kernels)
// peer_bases_dev / ctrl_off / data_off / region_bytes / channel: as moma_peer_exchange.  n = world * rows_per_rank query
// rows (ordered by owner rank); every rank calls merge_push, then combine_poll, on the same channel, once per step.
extern "C" __attribute__((visibility("default"))) int moma_nce_merge_push(
    const float* part_m, const float* part_l, const float* part_mmax, const float* part_O, int n_parts, int64_t rows_per_rank,
    int64_t D, const uint64_t* peer_bases_dev, int64_t ctrl_off, int64_t data_off, int64_t region_bytes, int rank, int world,
    int channel, moma_stream_t stream) {
    MOMA_REQUIRE(rows_per_rank > 0 && D > 0 && n_parts > 0, MOMA_ERR_INVALID, "nce_merge_push: bad shape");
    MOMA_REQUIRE(part_m && part_l && part_mmax && part_O && peer_bases_dev, MOMA_ERR_INVALID, "nce_merge_push: null pointer");
    MOMA_REQUIRE(D % 4 == 0 && D <= kCombMaxD && aligned16(part_O), MOMA_ERR_ALIGN, "nce_merge_push: D %% 4 != 0, D > %d or part_O unaligned", kCombMaxD);
    MOMA_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world && channel >= 0 && channel < kPeerChannels,
                 MOMA_ERR_INVALID, "nce_merge_push: bad rank / world / channel");
    MOMA_REQUIRE(2 * rows_per_rank * (D + 4) * 4 * world <= region_bytes && region_bytes % 16 == 0 && data_off % 16 == 0,
                 MOMA_ERR_WORKSPACE, "nce_merge_push: receive region too small");
    const int64_t B = rows_per_rank * world;
    const PeerLink link{reinterpret_cast<const unsigned long long*>(peer_bases_dev), (long long)ctrl_off, (long long)data_off,
                        (long long)region_bytes, rank, world, channel, (int)rows_per_rank};
    launch_pdl(nce_reduce_kernel<false, true>, dim3((unsigned)B), dim3(kCombThreads), 0, as_stream(stream),
        part_m, part_l, part_mmax, part_O, n_parts, nullptr, nullptr, (int)B, (int)D, 1.f, 0, 1.f,
        B, 1, B * D, D, 1, D, nullptr, nullptr, nullptr, nullptr, nullptr, link);
    MOMA_CUDA_LAUNCH_CHECK("nce_merge_push");
    note_launches(1);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_nce_combine_poll(
    const float* q_f32, const float* kpos_f32, int64_t B, int64_t D, float inv_T, int round_bf16, float dq_scale,
    const uint64_t* peer_bases_dev, int64_t ctrl_off, int64_t data_off, int64_t region_bytes, int rank, int world, int channel,
    float* loss_rows, float* dq_unit, int32_t* pos_is_max, float* max_logit, float* loss_mean, float* acc_pct,
    moma_stream_t stream) {
    MOMA_REQUIRE(B > 0 && D > 0, MOMA_ERR_INVALID, "nce_combine_poll: bad shape");
    MOMA_REQUIRE(q_f32 && kpos_f32 && loss_rows && dq_unit && pos_is_max && peer_bases_dev, MOMA_ERR_INVALID, "nce_combine_poll: null pointer");
    MOMA_REQUIRE(D % 4 == 0 && D <= kCombMaxD, MOMA_ERR_ALIGN, "nce_combine_poll: D %% 4 != 0 or D > %d", kCombMaxD);
    MOMA_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world && channel >= 0 && channel < kPeerChannels,
                 MOMA_ERR_INVALID, "nce_combine_poll: bad rank / world / channel");
    const PeerLink link{reinterpret_cast<const unsigned long long*>(peer_bases_dev), (long long)ctrl_off, (long long)data_off,
                        (long long)region_bytes, rank, world, channel, (int)B};
    const size_t smem = (size_t)world * (D + 4) * sizeof(float);
    launch_pdl(nce_reduce_kernel<true, true>, dim3((unsigned)B), dim3(kCombThreads), smem, as_stream(stream),
        nullptr, nullptr, nullptr, nullptr, world, q_f32, kpos_f32, (int)B, (int)D, inv_T, round_bf16, dq_scale,
        0, 0, 0, 0, 1, D, loss_rows, dq_unit, pos_is_max, max_logit, nullptr, link);
    MOMA_CUDA_LAUNCH_CHECK("nce_combine_poll");
    note_launches(1);
    if (loss_mean && acc_pct) {
        launch_pdl(nce_finalize_kernel, dim3(1), dim3(256), 0, as_stream(stream), loss_rows, pos_is_max, (int)B, loss_mean, acc_pct);
        MOMA_CUDA_LAUNCH_CHECK("nce_finalize");
        note_launches(1);
    }
    return MOMA_OK;
}

// ---- one launch for the whole pass: tcgen05 partial kernel + in-kernel combine (D = 64 / 128, bf16 operands)
extern "C" __attribute__((visibility("default"))) int moma_nce_fused_supported(int64_t B, int64_t D, int64_t K_local) {
    return nce_fused_supported(B, D, K_local) ? 1 : 0;
}
extern "C" __attribute__((visibility("default"))) size_t moma_nce_fused_workspace_bytes(int64_t B, int64_t D, int64_t K_local) {
    if (!nce_fused_supported(B, D, K_local)) return 0;
    return nce_fused_workspace_floats(B, D, K_local) * sizeof(float);
}
extern "C" __attribute__((visibility("default"))) int moma_nce_fused(
    const void* q_bf16, const void* queue_bf16, const float* q_f32, const float* kpos_f32, int64_t B, int64_t D,
    int64_t K_local, float inv_T, int round_bf16, float dq_scale, void* workspace, size_t workspace_bytes,
    uint32_t* counters, float* loss_rows, float* dq_unit, int32_t* pos_is_max, float* max_logit, float* loss_mean,
    float* acc_pct, moma_stream_t stream) {
    MOMA_REQUIRE(nce_fused_supported(B, D, K_local), MOMA_ERR_UNSUPPORTED, "nce_fused: unsupported shape B=%lld D=%lld K=%lld",
                 (long long)B, (long long)D, (long long)K_local);
    MOMA_REQUIRE(q_bf16 && queue_bf16 && q_f32 && kpos_f32 && workspace && counters && loss_rows && dq_unit && pos_is_max,
                 MOMA_ERR_INVALID, "nce_fused: null pointer");
    MOMA_REQUIRE((loss_mean == nullptr) == (acc_pct == nullptr), MOMA_ERR_INVALID, "nce_fused: loss_mean and acc_pct go together");
    MOMA_REQUIRE(workspace_bytes >= moma_nce_fused_workspace_bytes(B, D, K_local) && aligned16(workspace), MOMA_ERR_WORKSPACE,
                 "nce_fused: workspace too small or unaligned");
    MOMA_REQUIRE(aligned16(q_f32) && aligned16(kpos_f32) && aligned16(dq_unit), MOMA_ERR_ALIGN, "nce_fused: unaligned pointer");
    MOMA_REQUIRE(inv_T > 0.f, MOMA_ERR_INVALID, "nce_fused: temperature must be positive");
    tc::NceFuse f{};
    f.mode = 1; f.q_f32 = q_f32; f.kpos_f32 = kpos_f32; f.inv_T = inv_T; f.dq_scale = dq_scale; f.round_bf16 = round_bf16;
    f.loss_rows = loss_rows; f.dq = dq_unit; f.pos_is_max = pos_is_max; f.max_logit = max_logit; f.loss_mean = loss_mean;
    f.acc_pct = acc_pct; f.packed = nullptr; f.counters = counters;
    return nce_fused_launch(q_bf16, queue_bf16, B, D, K_local, inv_T, static_cast<float*>(workspace), f, as_stream(stream));
}
extern "C" __attribute__((visibility("default"))) int moma_nce_fused_packed(
    const void* q_bf16, const void* queue_bf16, int64_t B, int64_t D, int64_t K_local, float inv_T, void* workspace,
    size_t workspace_bytes, uint32_t* counters, float* packed, moma_stream_t stream) {
    MOMA_REQUIRE(nce_fused_supported(B, D, K_local), MOMA_ERR_UNSUPPORTED, "nce_fused_packed: unsupported shape");
    MOMA_REQUIRE(q_bf16 && queue_bf16 && workspace && counters && packed, MOMA_ERR_INVALID, "nce_fused_packed: null pointer");
    MOMA_REQUIRE(workspace_bytes >= moma_nce_fused_workspace_bytes(B, D, K_local) && aligned16(workspace) && aligned16(packed),
                 MOMA_ERR_WORKSPACE, "nce_fused_packed: workspace too small or unaligned");
    MOMA_REQUIRE(inv_T > 0.f, MOMA_ERR_INVALID, "nce_fused_packed: temperature must be positive");
    tc::NceFuse f{};
    f.mode = 2; f.inv_T = inv_T; f.dq_scale = 1.f; f.packed = packed; f.counters = counters;
    return nce_fused_launch(q_bf16, queue_bf16, B, D, K_local, inv_T, static_cast<float*>(workspace), f, as_stream(stream));
}

extern "C" __attribute__((visibility("default"))) int moma_nce_logits(const void* q, const void* kpos, const void* queue, int64_t B,
                               int64_t D, int64_t K, float T, int dtype, float* logits,
                               moma_stream_t stream) {
    MOMA_REQUIRE(B > 0 && D > 0 && K > 0 && q && kpos && queue && logits, MOMA_ERR_INVALID, "nce_logits: bad arguments");
    MOMA_REQUIRE(T != 0.f, MOMA_ERR_INVALID, "nce_logits: T == 0");
    const dim3 grid((unsigned)((K + 63) / 64), (unsigned)((B + 63) / 64));
    cudaStream_t st = as_stream(stream);
    if (dtype == MOMA_F32) {
        nce_logits_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(q), static_cast<const float*>(queue), (int)B, (int)D, K, T, logits);
        nce_pos_kernel<float><<<(unsigned)((B + 3) / 4), 128, 0, st>>>(static_cast<const float*>(q), static_cast<const float*>(kpos), (int)B, (int)D, T, logits, K + 1);
    } else if (dtype == MOMA_BF16) {
        using bf = __nv_bfloat16;
        nce_logits_kernel<bf><<<grid, 256, 0, st>>>(static_cast<const bf*>(q), static_cast<const bf*>(queue), (int)B, (int)D, K, T, logits);
        nce_pos_kernel<bf><<<(unsigned)((B + 3) / 4), 128, 0, st>>>(static_cast<const bf*>(q), static_cast<const bf*>(kpos), (int)B, (int)D, T, logits, K + 1);
    } else {
        return fail(MOMA_ERR_INVALID, "nce_logits: unknown dtype %d", dtype);
    }
    MOMA_CUDA_LAUNCH_CHECK("nce_logits");
    note_launches(2);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_nce_logits_qk(const float* q, const float* kpos, int64_t B, int64_t D, float T,
                                  float* out, moma_stream_t stream) {
    MOMA_REQUIRE(B > 0 && D > 0 && q && kpos && out && T != 0.f, MOMA_ERR_INVALID, "nce_logits_qk: bad arguments");
    nce_pos_kernel<float><<<(unsigned)((B + 3) / 4), 128, 0, as_stream(stream)>>>(q, kpos, (int)B, (int)D, T, out, 1);
    MOMA_CUDA_LAUNCH_CHECK("nce_logits_qk");
    note_launches(1);
    return MOMA_OK;
}
