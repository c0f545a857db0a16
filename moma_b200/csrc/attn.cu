// Multi-head attention over the batch axis (student/teacher embeddings as tokens).
// Replaces MoMA/criterion_moco_att.py:153-167 (Attention.forward) and its autograd
// backward.  FP32; projections are shared-memory tiled GEMMs, the score / softmax /
// value product is one fused flash-style kernel (shared-memory staged tiles,
// warp-shuffle row reductions), so the [H, N, N] score tensor never reaches HBM.
//
//   x [N, C] -> qkv = x W_qkv^T + b  ([N, 3C]; column = which*C + head*hd + d,
//   matching reshape(B, N, 3, H, hd) at :157) -> per head softmax(q k^T * hd^-0.5) v
//   -> o [N, C] (heads merged, :164) -> y = o W_proj^T + b_proj (:165).
#include <math_constants.h>
#include <mutex>
#include "common.cuh"

namespace moma {

// attn_tc.cu: the attention core on the tensor cores (3xTF32 mma.sync) for head_dim 8 / 16 / 32
bool attn_tc_supported(int hd);
void attn_tc_fwd(const float* qkv, int N, int C, int H, float scale, float* o, float* lse, cudaStream_t st, int q_start,
                 int q_stride, int NQ);
void attn_tc_bwd(const float* qkv, const float* dO, const float* lse, const float* delta, int N, int C, int H, float scale,
                 float* dqkv, cudaStream_t st, cudaStream_t st2);

// ------------------------------------------------------------------ generic SGEMM
// C[m, n] = sum_k A(m,k) * B(n,k) (+ bias[n]);  A(m,k) = A[m*a_rs + k*a_cs], same for B.
// The GEMMs of this path are tiny (M, N, K in the hundreds) and latency-bound, so the kernel is
// built for few dependent phases and many CTAs: 32x32 output tile per 128-thread CTA, the whole
// K strip (up to 128) staged in shared memory in ONE load phase, 128-bit shared-memory reads.
constexpr int kGN = 32, kGK = 128, kGLd = kGK + 4, kGThreads = 128;

template <bool KCONTIG, int ROWS>
__device__ __forceinline__ void load_strip(float* dst, const float* __restrict__ src, int64_t rs, int64_t cs,
                                           int r0, int R, int k0, int K, bool vec_ok) {
    const int tid = threadIdx.x;
    if (KCONTIG) {
        // consecutive lanes walk along k (contiguous): float4 when aligned
#pragma unroll
        for (int i = 0; i < (ROWS * kGK / 4) / kGThreads; ++i) {
            const int idx = tid + i * kGThreads;
            const int r = idx >> 5, v = idx & 31;
            const int k = k0 + 4 * v;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < R) {
                const float* p = src + (int64_t)(r0 + r) * rs + k;
                if (vec_ok && k + 3 < K) val = *reinterpret_cast<const float4*>(p);
                else {
                    if (k < K) val.x = p[0];
                    if (k + 1 < K) val.y = p[1];
                    if (k + 2 < K) val.z = p[2];
                    if (k + 3 < K) val.w = p[3];
                }
            }
            *reinterpret_cast<float4*>(dst + r * kGLd + 4 * v) = val;
        }
    } else {
        // consecutive lanes walk along the row index (contiguous in memory), k strided by cs
#pragma unroll 8
        for (int i = 0; i < (ROWS * kGK) / kGThreads; ++i) {
            const int idx = tid + i * kGThreads;
            const int r = idx % ROWS, k = idx / ROWS;
            dst[r * kGLd + k] = (r0 + r < R && k0 + k < K) ? src[(int64_t)(k0 + k) * cs + (r0 + r)] : 0.f;
        }
    }
}

// TM = 32 or 16 output rows per CTA (16 when the 32-row grid would leave most SMs idle)
template <bool A_KCONTIG, bool B_KCONTIG, int TM>
__global__ void __launch_bounds__(kGThreads)
sgemm_kernel(const float* __restrict__ A, int64_t a_rs, int64_t a_cs, const float* __restrict__ Bm,
             int64_t b_rs, int64_t b_cs, const float* __restrict__ bias, float* __restrict__ Cm,
             int64_t ldc, int M, int N, int K, int a_vec, int b_vec) {
    __shared__ __align__(16) float As[TM * kGLd];
    __shared__ __align__(16) float Bs[kGN * kGLd];
    constexpr int RT = TM / 16;                                  // rows per thread
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * kGN;
    float acc[RT][4] = {};
    for (int k0 = 0; k0 < K; k0 += kGK) {
        if (k0 > 0) __syncthreads();
        load_strip<A_KCONTIG, TM>(As, A, a_rs, a_cs, m0, M, k0, K, a_vec != 0);
        load_strip<B_KCONTIG, kGN>(Bs, Bm, b_rs, b_cs, n0, N, k0, K, b_vec != 0);
        __syncthreads();
        const int kend = min(kGK, K - k0);
#pragma unroll 4
        for (int k = 0; k < kend; k += 4) {
            float4 a[RT], b[4];
#pragma unroll
            for (int i = 0; i < RT; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * kGLd + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(Bs + (tx + 8 * j) * kGLd + k);
#pragma unroll
            for (int i = 0; i < RT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[i][j] += a[i].x * b[j].x + a[i].y * b[j].y + a[i].z * b[j].z + a[i].w * b[j].w;
        }
    }
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int r = m0 + ty + 16 * i;
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx + 8 * j;
            if (c < N) Cm[(int64_t)r * ldc + c] = acc[i][j] + (bias ? bias[c] : 0.f);
        }
    }
}

template <int TM>
static void sgemm_tm(const float* A, int64_t a_rs, int64_t a_cs, const float* Bm, int64_t b_rs,
                     int64_t b_cs, const float* bias, float* Cm, int64_t ldc, int M, int N, int K, cudaStream_t st) {
    const dim3 grid((N + kGN - 1) / kGN, (M + TM - 1) / TM);
    const bool ak = (a_cs == 1), bk = (b_cs == 1);
    const int av = ak && (a_rs % 4 == 0) && aligned16(A), bv = bk && (b_rs % 4 == 0) && aligned16(Bm);
    if (ak && bk) sgemm_kernel<true, true, TM><<<grid, kGThreads, 0, st>>>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K, av, bv);
    else if (ak && !bk) sgemm_kernel<true, false, TM><<<grid, kGThreads, 0, st>>>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K, av, bv);
    else if (!ak && bk) sgemm_kernel<false, true, TM><<<grid, kGThreads, 0, st>>>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K, av, bv);
    else sgemm_kernel<false, false, TM><<<grid, kGThreads, 0, st>>>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K, av, bv);
}

static void sgemm(const float* A, int64_t a_rs, int64_t a_cs, const float* Bm, int64_t b_rs,
                  int64_t b_cs, const float* bias, float* Cm, int64_t ldc, int M, int N, int K,
                  cudaStream_t st, void* c_bf16 = nullptr) {
    if (!use_simt_gemm()) {      // tensor-core path (3xTF32, gemm.cu); single split: no workspace here
        gemm_nt(A, nullptr, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K, 0, nullptr, 0, st, c_bf16);
        return;
    }
    const int ctas32 = ((N + kGN - 1) / kGN) * ((M + 31) / 32);
    if (ctas32 >= sm_count() * 3 / 4) sgemm_tm<32>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K, st);
    else sgemm_tm<16>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K, st);
}


// ------------------------------------------------------------------ fused attention core
// Tiles: 4*RW "row" tokens per CTA (RW per warp, 4 warps) against KT "column" tokens per step.  KT is
// as wide as shared memory allows (256 for head_dim <= 32) so that at the path's sizes (N = 256..1024)
// a CTA goes through 1-4 load phases instead of N/64, and RW (8, 4 or 2) is picked at launch so the
// grid has about one CTA per SM: these kernels are latency-bound, not FLOP-bound.
constexpr int kAThreads = 128;

template <int HD, int RW> struct AttnCfg {
    static constexpr int AR = 4 * RW;                       // rows per CTA
    static constexpr int LD = HD + 4;                       // padded smem row (floats)
    static constexpr int CPL = HD >= 32 ? HD / 32 : 1;      // output columns per lane
    static constexpr int RPL = HD >= 32 ? RW : RW * HD / 32;   // output rows per lane (32/HD row groups)
    static_assert(RW * HD >= 32, "too few rows per warp for this head_dim");
    static constexpr int KT_FWD = HD <= 32 ? 256 : (HD == 64 ? 128 : 64);   // column tile, forward
    static constexpr int KT_BWD = HD <= 32 ? 128 : 64;                      // column tile, backward (2 score tiles live)
    __device__ static int roff(int lane) { return HD >= 32 ? 0 : (lane / HD) * RPL; }
    __device__ static int cbase(int lane) { return HD >= 32 ? lane : (lane % HD); }
};

// load `rows` token rows x HD columns (global row stride ldg) into smem [rows][HD+4]
template <int HD>
__device__ __forceinline__ void load_rows(float* dst, const float* __restrict__ src, int64_t ldg,
                                          int row0, int rows, int N) {
    constexpr int NV = HD / 4, LD = HD + 4;
    for (int i = threadIdx.x; i < rows * NV; i += kAThreads) {
        const int r = i / NV, v = i - r * NV;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < N) a = *reinterpret_cast<const float4*>(src + (int64_t)(row0 + r) * ldg + 4 * v);
        *reinterpret_cast<float4*>(dst + r * LD + 4 * v) = a;
    }
}

// same, asynchronously (cp.async, zero-filled past row N): the caller commits / waits the group
template <int HD>
__device__ __forceinline__ void load_rows_async(float* dst, const float* __restrict__ src, int64_t ldg,
                                                int row0, int rows, int N) {
    constexpr int NV = HD / 4, LD = HD + 4;
    for (int i = threadIdx.x; i < rows * NV; i += kAThreads) {
        const int r = i / NV, v = i - r * NV;
        const bool ok = row0 + r < N;
        cp_async16(dst + r * LD + 4 * v, ok ? src + (int64_t)(row0 + r) * ldg + 4 * v : src, ok);
    }
}

// s[r][cc] = sum_d a[warp*RW + r][d] * b[lane + 32*cc][d],  cc < KT/32
template <int HD, int KT, int RW>
__device__ __forceinline__ void dot_tile(const float* a_s, const float* b_s, int warp, int lane,
                                         float (&s)[RW][KT / 32]) {
    constexpr int LD = HD + 4, CC = KT / 32;
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) s[r][cc] = 0.f;
#pragma unroll 2
    for (int v = 0; v < HD / 4; ++v) {
        float4 b[CC];
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) b[cc] = *reinterpret_cast<const float4*>(b_s + (lane + 32 * cc) * LD + 4 * v);
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(a_s + (warp * RW + r) * LD + 4 * v);
#pragma unroll
            for (int cc = 0; cc < CC; ++cc)
                s[r][cc] += a.x * b[cc].x + a.y * b[cc].y + a.z * b[cc].z + a.w * b[cc].w;
        }
    }
}

// acc[r][c] += sum_j p[row(r)][j] * v[j][col(c)],  j over the KT-column tile
template <int HD, int KT, int RW>
__device__ __forceinline__ void acc_tile(const float* p_s, const float* v_s, int warp, int lane,
                                         float (&acc)[AttnCfg<HD, RW>::RPL][AttnCfg<HD, RW>::CPL]) {
    using Cfg = AttnCfg<HD, RW>;
    constexpr int LD = Cfg::LD, LDP = KT + 4;
    const int rbase = warp * RW + Cfg::roff(lane);
    const int cbase = Cfg::cbase(lane);
#pragma unroll 2
    for (int j = 0; j < KT; j += 4) {
        float vv[4][Cfg::CPL];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c) vv[jj][c] = v_s[(j + jj) * LD + cbase + 32 * c];
#pragma unroll
        for (int r = 0; r < Cfg::RPL; ++r) {
            const float4 p = *reinterpret_cast<const float4*>(p_s + (rbase + r) * LDP + j);
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c)
                acc[r][c] += p.x * vv[0][c] + p.y * vv[1][c] + p.z * vv[2][c] + p.w * vv[3][c];
        }
    }
}

// ---- forward: grid (ceil(N/32), H)
template <int HD, int RW>
__global__ void __launch_bounds__(kAThreads)
attn_fwd_kernel(const float* __restrict__ qkv, int N, int C, float scale, float* __restrict__ o,
                float* __restrict__ lse, int q_start, int q_stride, int NQ) {
    pdl_wait();
    pdl_launch_dependents();
    // queries are the NQ rows q_start + i * q_stride of the N tokens (all rows by default); o / lse are
    // indexed by the compact query index i
    using Cfg = AttnCfg<HD, RW>;
    constexpr int kAR = Cfg::AR;
    constexpr int KT = Cfg::KT_FWD, CC = KT / 32, LDP = KT + 4;
    extern __shared__ __align__(16) float sm[];
    float* q_s = sm;                         // [32][LD]
    float* k_s = q_s + kAR * Cfg::LD;        // [KT][LD]
    float* v_s = k_s + KT * Cfg::LD;         // [KT][LD]
    float* p_s = v_s + KT * Cfg::LD;         // [32][LDP]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.y, i0 = blockIdx.x * kAR;
    const int64_t ldg = 3 * (int64_t)C;
    const float* qg = qkv + h * HD;
    const float* kg = qkv + C + h * HD;
    const float* vg = qkv + 2 * C + h * HD;

    {
        constexpr int NV = HD / 4;
        for (int i = threadIdx.x; i < kAR * NV; i += kAThreads) {
            const int r = i / NV, v = i - r * NV;
            const bool ok = i0 + r < NQ;
            cp_async16(q_s + r * Cfg::LD + 4 * v, ok ? qg + (int64_t)(q_start + (i0 + r) * q_stride) * ldg + 4 * v : qg, ok);
        }
    }
    float m_run[RW], l_run[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) { m_run[r] = -CUDART_INF_F; l_run[r] = 0.f; }
    float acc[Cfg::RPL][Cfg::CPL] = {};
    const int roff = Cfg::roff(lane);

    for (int j0 = 0; j0 < N; j0 += KT) {
        if (j0 > 0) __syncthreads();
        // K (with Q on the first tile) and V arrive as two cp.async groups: the scores start as soon as K is in,
        // the V tile lands while they and the softmax are computed
        load_rows_async<HD>(k_s, kg, ldg, j0, KT, N);
        cp_async_commit();
        load_rows_async<HD>(v_s, vg, ldg, j0, KT, N);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        float s[RW][CC];
        dot_tile<HD, KT, RW>(q_s, k_s, warp, lane, s);
        float corr[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            float mx = -CUDART_INF_F;
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
                s[r][cc] = (j0 + lane + 32 * cc < N) ? s[r][cc] * scale : -CUDART_INF_F;
                mx = fmaxf(mx, s[r][cc]);
            }
            mx = warp_max(mx);
            const float m_new = fmaxf(m_run[r], mx);
            corr[r] = (m_run[r] == -CUDART_INF_F) ? 0.f : expf(m_run[r] - m_new);
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
                const float p = (s[r][cc] == -CUDART_INF_F) ? 0.f : expf(s[r][cc] - m_new);
                p_s[(warp * RW + r) * LDP + lane + 32 * cc] = p;
                sum += p;
            }
            l_run[r] = l_run[r] * corr[r] + warp_sum(sum);
            m_run[r] = m_new;
        }
#pragma unroll
        for (int r = 0; r < Cfg::RPL; ++r) {
            float cr = corr[0];                 // corr is warp-uniform per row; pick this lane's rows
#pragma unroll
            for (int rr = 0; rr < RW; ++rr) cr = (rr == r + roff) ? corr[rr] : cr;
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c) acc[r][c] *= cr;
        }
        cp_async_wait<0>();
        __syncthreads();
        acc_tile<HD, KT, RW>(p_s, v_s, warp, lane, acc);
    }
    const int cbase = Cfg::cbase(lane);
#pragma unroll
    for (int r = 0; r < Cfg::RPL; ++r) {
        float lr = l_run[0];
#pragma unroll
        for (int rr = 0; rr < RW; ++rr) lr = (rr == r + roff) ? l_run[rr] : lr;
        const int row = i0 + warp * RW + roff + r;
        if (row < NQ) {
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c)
                o[(int64_t)row * C + h * HD + cbase + 32 * c] = acc[r][c] / lr;
        }
    }
    if (lane < RW) {
        float mr = m_run[0], lr = l_run[0];
#pragma unroll
        for (int rr = 0; rr < RW; ++rr) { mr = (rr == lane) ? m_run[rr] : mr; lr = (rr == lane) ? l_run[rr] : lr; }
        const int row = i0 + warp * RW + lane;
        if (row < NQ) lse[(int64_t)h * NQ + row] = mr + logf(lr);
    }
}

// attention probabilities for Attention_viz: probs[h, i, j] = exp(s_ij - lse_i)
__global__ void __launch_bounds__(256)
attn_probs_kernel(const float* __restrict__ qkv, const float* __restrict__ lse, int N, int C, int H,
                  float scale, float* __restrict__ probs) {
    pdl_wait();
    pdl_launch_dependents();
    const int hd = C / H;
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (int64_t)H * N * N) return;
    const int j = (int)(idx % N);
    const int i = (int)((idx / N) % N);
    const int h = (int)(idx / ((int64_t)N * N));
    const float* q = qkv + (int64_t)i * 3 * C + h * hd;
    const float* k = qkv + (int64_t)j * 3 * C + C + h * hd;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s += q[d] * k[d];
    probs[idx] = expf(s * scale - lse[(int64_t)h * N + i]);
}

// ---- backward helpers
// delta[h, i] = sum_d do[i, h*hd + d] * o[i, h*hd + d]; one warp per (i, h)
__global__ void __launch_bounds__(128)
attn_delta_kernel(const float* __restrict__ dO, const float* __restrict__ o, int N, int C, int H,
                  float* __restrict__ delta) {
    pdl_wait();
    pdl_launch_dependents();
    const int w = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= N * H) return;
    const int i = w / H, h = w - i * H, hd = C / H;
    float s = 0.f;
    for (int d = lane; d < hd; d += 32)
        s += dO[(int64_t)i * C + h * hd + d] * o[(int64_t)i * C + h * hd + d];
    s = warp_sum(s);
    if (lane == 0) delta[(int64_t)h * N + i] = s;
}

// dQ: grid (ceil(N/32) query blocks, H)
template <int HD, int RW>
__global__ void __launch_bounds__(kAThreads)
attn_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ dO,
                   const float* __restrict__ lse, const float* __restrict__ delta, int N, int C,
                   float scale, float* __restrict__ dqkv) {
    pdl_wait();
    pdl_launch_dependents();
    using Cfg = AttnCfg<HD, RW>;
    constexpr int kAR = Cfg::AR;
    constexpr int KT = Cfg::KT_BWD, CC = KT / 32, LDP = KT + 4;
    extern __shared__ __align__(16) float sm[];
    float* q_s = sm;                          // [32][LD]
    float* do_s = q_s + kAR * Cfg::LD;        // [32][LD]
    float* kv_s = do_s + kAR * Cfg::LD;       // 2 x { [KT][LD] keys, [KT][LD] values }: double-buffered column tiles
    float* p_s = kv_s + 4 * KT * Cfg::LD;     // [32][LDP]  (holds dS)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.y, i0 = blockIdx.x * kAR;
    const int64_t ldg = 3 * (int64_t)C;
    auto issue_kv = [&](int buf, int j0) {
        load_rows_async<HD>(kv_s + (2 * buf) * KT * Cfg::LD, qkv + C + h * HD, ldg, j0, KT, N);
        load_rows_async<HD>(kv_s + (2 * buf + 1) * KT * Cfg::LD, qkv + 2 * C + h * HD, ldg, j0, KT, N);
        cp_async_commit();
    };
    load_rows_async<HD>(q_s, qkv + h * HD, ldg, i0, kAR, N);
    load_rows_async<HD>(do_s, dO + h * HD, C, i0, kAR, N);
    issue_kv(0, 0);
    float lse_r[RW], del_r[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        const int row = i0 + warp * RW + r;
        lse_r[r] = row < N ? lse[(int64_t)h * N + row] : 0.f;
        del_r[r] = row < N ? delta[(int64_t)h * N + row] : 0.f;
    }
    float acc[Cfg::RPL][Cfg::CPL] = {};
    for (int j0 = 0, it = 0; j0 < N; j0 += KT, ++it) {
        cp_async_wait<0>();
        __syncthreads();                      // tile `it` has landed; everyone is done with the other buffer
        if (j0 + KT < N) issue_kv((it + 1) & 1, j0 + KT);       // next tile in flight during this one's math
        const float* k_s = kv_s + (2 * (it & 1)) * KT * Cfg::LD;
        const float* v_s = k_s + KT * Cfg::LD;
        float s[RW][CC], dp[RW][CC];
        dot_tile<HD, KT, RW>(q_s, k_s, warp, lane, s);
        dot_tile<HD, KT, RW>(do_s, v_s, warp, lane, dp);
#pragma unroll
        for (int r = 0; r < RW; ++r) {
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
                const bool ok = (j0 + lane + 32 * cc < N);
                const float p = ok ? expf(s[r][cc] * scale - lse_r[r]) : 0.f;
                p_s[(warp * RW + r) * LDP + lane + 32 * cc] = p * (dp[r][cc] - del_r[r]);
            }
        }
        __syncwarp();
        acc_tile<HD, KT, RW>(p_s, k_s, warp, lane, acc);
    }
    const int roff = Cfg::roff(lane);
    const int cbase = Cfg::cbase(lane);
#pragma unroll
    for (int r = 0; r < Cfg::RPL; ++r) {
        const int row = i0 + warp * RW + roff + r;
        if (row < N)
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c)
                dqkv[(int64_t)row * ldg + h * HD + cbase + 32 * c] = acc[r][c] * scale;
    }
}

// dK, dV: grid (ceil(N/32) key blocks, H); rows = keys, columns = queries
template <int HD, int RW>
__global__ void __launch_bounds__(kAThreads)
attn_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ dO,
                    const float* __restrict__ lse, const float* __restrict__ delta, int N, int C,
                    float scale, float* __restrict__ dqkv) {
    pdl_wait();
    pdl_launch_dependents();
    using Cfg = AttnCfg<HD, RW>;
    constexpr int kAR = Cfg::AR;
    constexpr int KT = Cfg::KT_BWD, CC = KT / 32, LDP = KT + 4;
    extern __shared__ __align__(16) float sm[];
    float* k_s = sm;                           // [32][LD]  own keys
    float* v_s = k_s + kAR * Cfg::LD;          // [32][LD]  own values
    float* qd_s = v_s + kAR * Cfg::LD;         // 2 x { [KT][LD] query tile, [KT][LD] dO tile }: double-buffered
    float* p_s = qd_s + 4 * KT * Cfg::LD;      // [32][LDP]  P^T
    float* ds_s = p_s + kAR * LDP;             // [32][LDP]  dS^T
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.y, j0 = blockIdx.x * kAR;
    const int64_t ldg = 3 * (int64_t)C;
    auto issue_qd = [&](int buf, int i0) {
        load_rows_async<HD>(qd_s + (2 * buf) * KT * Cfg::LD, qkv + h * HD, ldg, i0, KT, N);
        load_rows_async<HD>(qd_s + (2 * buf + 1) * KT * Cfg::LD, dO + h * HD, C, i0, KT, N);
        cp_async_commit();
    };
    load_rows_async<HD>(k_s, qkv + C + h * HD, ldg, j0, kAR, N);
    load_rows_async<HD>(v_s, qkv + 2 * C + h * HD, ldg, j0, kAR, N);
    issue_qd(0, (int)((long long)((N + KT - 1) / KT) * blockIdx.z / gridDim.z) * KT);
    float acc_k[Cfg::RPL][Cfg::CPL] = {};
    float acc_v[Cfg::RPL][Cfg::CPL] = {};
    // gridDim.z == 2: the query tiles are split between two CTAs whose results are added into a zeroed dK/dV
    // (exactly two addends per element: the sum does not depend on their order -> still deterministic)
    const int n_tiles = (N + KT - 1) / KT;
    const int t_begin = (int)((long long)n_tiles * blockIdx.z / gridDim.z), t_end = (int)((long long)n_tiles * (blockIdx.z + 1) / gridDim.z);
    const int i_end = min(N, t_end * KT);
    for (int i0 = t_begin * KT, it = 0; i0 < i_end; i0 += KT, ++it) {
        cp_async_wait<0>();
        __syncthreads();
        if (i0 + KT < i_end) issue_qd((it + 1) & 1, i0 + KT);
        const float* q_s = qd_s + (2 * (it & 1)) * KT * Cfg::LD;
        const float* do_s = q_s + KT * Cfg::LD;
        float lse_c[CC], del_c[CC];                          // issued before the dot products: their latency hides there
#pragma unroll
        for (int cc = 0; cc < CC; ++cc) {
            const int i = i0 + lane + 32 * cc;
            lse_c[cc] = i < N ? lse[(int64_t)h * N + i] : 0.f;
            del_c[cc] = i < N ? delta[(int64_t)h * N + i] : 0.f;
        }
        float st[RW][CC], dpt[RW][CC];
        dot_tile<HD, KT, RW>(k_s, q_s, warp, lane, st);      // st[r][cc] = k_j . q_i
        dot_tile<HD, KT, RW>(v_s, do_s, warp, lane, dpt);    // dpt      = v_j . do_i
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const bool rok = (j0 + warp * RW + r < N);
#pragma unroll
            for (int cc = 0; cc < CC; ++cc) {
                const bool ok = rok && (i0 + lane + 32 * cc < N);
                const float p = ok ? expf(st[r][cc] * scale - lse_c[cc]) : 0.f;
                p_s[(warp * RW + r) * LDP + lane + 32 * cc] = p;
                ds_s[(warp * RW + r) * LDP + lane + 32 * cc] = p * (dpt[r][cc] - del_c[cc]);
            }
        }
        __syncwarp();
        acc_tile<HD, KT, RW>(p_s, do_s, warp, lane, acc_v);
        acc_tile<HD, KT, RW>(ds_s, q_s, warp, lane, acc_k);
    }
    const int roff = Cfg::roff(lane);
    const int cbase = Cfg::cbase(lane);
#pragma unroll
    for (int r = 0; r < Cfg::RPL; ++r) {
        const int row = j0 + warp * RW + roff + r;
        if (row < N)
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c) {
                float* pk = dqkv + (int64_t)row * ldg + C + h * HD + cbase + 32 * c;
                float* pv = dqkv + (int64_t)row * ldg + 2 * C + h * HD + cbase + 32 * c;
                if (gridDim.z == 1) { *pk = acc_k[r][c] * scale; *pv = acc_v[r][c]; }
                else { atomicAdd(pk, acc_k[r][c] * scale); atomicAdd(pv, acc_v[r][c]); }
            }
    }
}

template <int HD, int RW> static size_t fwd_smem() {
    constexpr int KT = AttnCfg<HD, RW>::KT_FWD, AR = 4 * RW;
    return (size_t)((AR + 2 * KT) * (HD + 4) + AR * (KT + 4)) * sizeof(float);
}
template <int HD, int RW> static size_t dq_smem() {
    constexpr int KT = AttnCfg<HD, RW>::KT_BWD, AR = 4 * RW;
    return (size_t)((2 * AR + 4 * KT) * (HD + 4) + AR * (KT + 4)) * sizeof(float);
}
template <int HD, int RW> static size_t dkv_smem() {
    constexpr int KT = AttnCfg<HD, RW>::KT_BWD, AR = 4 * RW;
    return (size_t)((2 * AR + 4 * KT) * (HD + 4) + 2 * AR * (KT + 4)) * sizeof(float);
}

template <int HD, int RW>
static void launch_fwd_rw(const float* qkv, int N, int C, int H, float scale, float* o, float* lse, cudaStream_t st,
                          int q_start, int q_stride, int NQ) {
    ensure_dyn_smem(reinterpret_cast<const void*>(attn_fwd_kernel<HD, RW>), (int)fwd_smem<HD, RW>());
    constexpr int AR = 4 * RW;
    launch_pdl(attn_fwd_kernel<HD, RW>, dim3((NQ + AR - 1) / AR, H), dim3(kAThreads), fwd_smem<HD, RW>(), st, qkv, N, C, scale, o, lse,
               q_start, q_stride, NQ);
}
template <int HD, int RW>
static void launch_bwd_rw(const float* qkv, const float* dO, const float* lse, const float* delta, int N, int C, int H,
                          float scale, float* dqkv, cudaStream_t st, cudaStream_t st2, bool dkv_split) {
    ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dq_kernel<HD, RW>), (int)dq_smem<HD, RW>());
    ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dkv_kernel<HD, RW>), (int)dkv_smem<HD, RW>());
    constexpr int AR = 4 * RW;
    const dim3 grid((N + AR - 1) / AR, H);
    // dQ and dK/dV are independent: two streams (st2 was forked from st by the caller)
    launch_pdl(attn_bwd_dq_kernel<HD, RW>, grid, dim3(kAThreads), dq_smem<HD, RW>(), st, qkv, dO, lse, delta, N, C, scale, dqkv);
    launch_pdl(attn_bwd_dkv_kernel<HD, RW>, dim3(grid.x, grid.y, dkv_split ? 2 : 1), dim3(kAThreads), dkv_smem<HD, RW>(), st2,
               qkv, dO, lse, delta, N, C, scale, dqkv);
}


// rows per warp: the largest of {8, 4, 2} (but RW * HD >= 32) that still gives about one CTA per SM.  (One row per
// warp -- two CTAs per SM -- was measured: each warp still reads the whole key tile from shared memory, so the
// instruction count per SM rises as fast as the occupancy and the kernels do not get faster.)
template <int HD> static int pick_rw(int N, int H) {
    constexpr int rw_min = HD >= 16 ? 2 : 4;
    const int target = sm_count() * 3 / 4;
    for (int rw = 8; rw > rw_min; rw >>= 1)
        if (((N + 4 * rw - 1) / (4 * rw)) * H >= target) return rw;
    return rw_min;
}
template <int HD>
static void launch_fwd(const float* qkv, int N, int C, int H, float scale, float* o, float* lse, cudaStream_t st,
                       int q_start = 0, int q_stride = 1, int NQ = -1) {
    if (NQ < 0) NQ = N;
    note_flops(1, 4.0 * NQ * N * C);                      // Q K^T and P V, 2 FLOP per MAC
    switch (pick_rw<HD>(NQ, H)) {
        case 8: launch_fwd_rw<HD, 8>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ); break;
        case 4: launch_fwd_rw<HD, 4>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ); break;
        default: if constexpr (HD >= 16) launch_fwd_rw<HD, 2>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ);
                 else launch_fwd_rw<HD, 4>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ);
    }
}
// dK/dV is the longest kernel of the backward (twice the work of dQ): split its query loop over two CTAs when
// there are at least two query tiles and the grid is at most one wave (the caller then zeroes dK/dV first)
template <int HD>
static bool dkv_wants_split(int N, int H) {
    const int rw = pick_rw<HD>(N, H);
    constexpr int KT = HD <= 32 ? 128 : 64;                 // AttnCfg::KT_BWD
    return (N + KT - 1) / KT >= 2 && ((N + 4 * rw - 1) / (4 * rw)) * H <= sm_count();
}
template <int HD>
static void launch_bwd(const float* qkv, const float* dO, const float* lse, const float* delta, int N, int C, int H,
                       float scale, float* dqkv, cudaStream_t st, cudaStream_t st2, bool dkv_split) {
    note_flops(1, 8.0 * N * N * C);                       // dV, dP, dQ, dK (the recomputed scores are not counted)
    switch (pick_rw<HD>(N, H)) {
        case 8: launch_bwd_rw<HD, 8>(qkv, dO, lse, delta, N, C, H, scale, dqkv, st, st2, dkv_split); break;
        case 4: launch_bwd_rw<HD, 4>(qkv, dO, lse, delta, N, C, H, scale, dqkv, st, st2, dkv_split); break;
        default: if constexpr (HD >= 16) launch_bwd_rw<HD, 2>(qkv, dO, lse, delta, N, C, H, scale, dqkv, st, st2, dkv_split);
                 else launch_bwd_rw<HD, 4>(qkv, dO, lse, delta, N, C, H, scale, dqkv, st, st2, dkv_split);
    }
}

// Side streams + events for the fork/join inside moma_attn_bwd (created once per process; no device memory).
struct BwdStreams {
    cudaStream_t s1 = nullptr, s2 = nullptr;
    cudaEvent_t fork = nullptr, d_o = nullptr, dq = nullptr, dkv = nullptr, join1 = nullptr, join2 = nullptr;
    bool ok = false;
};
static BwdStreams& bwd_streams() {
    // one set per device, created once under std::call_once (streams and events belong to a device)
    constexpr int kMaxDev = 64;
    static BwdStreams per_dev[kMaxDev];
    static std::once_flag once[kMaxDev];
    static BwdStreams none;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { cudaGetLastError(); return none; }
    BwdStreams& b = per_dev[dev];
    std::call_once(once[dev], [&b] {
        bool good = cudaStreamCreateWithFlags(&b.s1, cudaStreamNonBlocking) == cudaSuccess &&
                    cudaStreamCreateWithFlags(&b.s2, cudaStreamNonBlocking) == cudaSuccess;
        cudaEvent_t* evs[] = {&b.fork, &b.d_o, &b.dq, &b.dkv, &b.join1, &b.join2};
        for (auto e : evs) good = good && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
        b.ok = good;
        if (!good) cudaGetLastError();
    });
    return b;
}

static int check_attn(const char* who, int64_t N, int64_t C, int H) {
    MOMA_REQUIRE(N > 0 && C > 0 && H > 0, MOMA_ERR_INVALID, "%s: bad shape N=%lld C=%lld H=%d", who, (long long)N, (long long)C, H);
    MOMA_REQUIRE(C % H == 0, MOMA_ERR_INVALID, "%s: C=%lld not divisible by H=%d", who, (long long)C, H);
    const int64_t hd = C / H;
    MOMA_REQUIRE(hd == 8 || hd == 16 || hd == 32 || hd == 64 || hd == 128, MOMA_ERR_UNSUPPORTED,
                 "%s: head_dim=%lld unsupported (8, 16, 32, 64, 128)", who, (long long)hd);
    MOMA_REQUIRE(N < (1 << 24) && C <= 8192, MOMA_ERR_UNSUPPORTED, "%s: shape too large", who);
    return MOMA_OK;
}

}  // namespace moma

using namespace moma;

static inline bool ldc_is_dense(int64_t ldc, int n) { return ldc == n; }

extern "C" __attribute__((visibility("default"))) int moma_attn_fwd(const float* x, const float* w_qkv, const float* b_qkv,
                             const float* w_proj, const float* b_proj, int64_t N, int64_t C, int H,
                             float* y, float* qkv, float* o, float* lse, float* attn_probs, void* y_bf16,
                             moma_stream_t stream) {
    int rc = check_attn("attn_fwd", N, C, H);
    if (rc != MOMA_OK) return rc;
    MOMA_REQUIRE(x && w_qkv && w_proj && b_proj && y && qkv && o && lse, MOMA_ERR_INVALID, "attn_fwd: null pointer");
    MOMA_REQUIRE(aligned16(qkv) && aligned16(o), MOMA_ERR_ALIGN, "attn_fwd: qkv/o must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int n = (int)N, c = (int)C, hd = c / H;
    const float scale = 1.0f / sqrtf((float)hd);
    sgemm(x, C, 1, w_qkv, C, 1, b_qkv, qkv, 3 * C, n, 3 * c, c, st);
    if (attn_tc_supported(hd)) attn_tc_fwd(qkv, n, c, H, scale, o, lse, st, 0, 1, n);
    else switch (hd) {
        case 8: launch_fwd<8>(qkv, n, c, H, scale, o, lse, st); break;
        case 16: launch_fwd<16>(qkv, n, c, H, scale, o, lse, st); break;
        case 32: launch_fwd<32>(qkv, n, c, H, scale, o, lse, st); break;
        case 64: launch_fwd<64>(qkv, n, c, H, scale, o, lse, st); break;
        default: launch_fwd<128>(qkv, n, c, H, scale, o, lse, st); break;
    }
    if (attn_probs) {
        const int64_t tot = (int64_t)H * N * N;
        launch_pdl(attn_probs_kernel, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, st, qkv, lse, n, c, H, scale, attn_probs);
    }
    const bool fused_copy = y_bf16 != nullptr && !use_simt_gemm() && ldc_is_dense(C, c);
    sgemm(o, C, 1, w_proj, C, 1, b_proj, y, C, n, c, c, st, fused_copy ? y_bf16 : nullptr);
    if (y_bf16 != nullptr && !fused_copy) {
        int rc2 = moma_cast_bf16(y, y_bf16, N * C, stream);
        if (rc2 != MOMA_OK) return rc2;
    }
    MOMA_CUDA_LAUNCH_CHECK("attn_fwd");
    note_launches(attn_probs ? 4 : 3);
    return MOMA_OK;
}

// Forward for a strided SUBSET of the query rows (rows q_start + i * q_stride, i < q_count): keys / values still
// come from all N tokens.  Used by the K-sharded queue, where a rank only needs the attended keys it will
// enqueue (every W-th row of the all-gathered keys): N * N / W score work instead of N * N.  No backward.
extern "C" __attribute__((visibility("default"))) int moma_attn_fwd_rows(
    const float* x, const float* w_qkv, const float* b_qkv, const float* w_proj, const float* b_proj, int64_t N,
    int64_t C, int H, int64_t q_start, int64_t q_stride, int64_t q_count, float* y, float* qkv, float* o,
    float* lse, moma_stream_t stream) {
    int rc = check_attn("attn_fwd_rows", N, C, H);
    if (rc != MOMA_OK) return rc;
    MOMA_REQUIRE((x == nullptr || w_qkv) && w_proj && b_proj && y && qkv && o && lse, MOMA_ERR_INVALID, "attn_fwd_rows: null pointer");
    MOMA_REQUIRE(q_count > 0 && q_stride > 0 && q_start >= 0 && q_start + (q_count - 1) * q_stride < N, MOMA_ERR_INVALID,
                 "attn_fwd_rows: row subset out of range");
    MOMA_REQUIRE(aligned16(qkv) && aligned16(o), MOMA_ERR_ALIGN, "attn_fwd_rows: qkv/o must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int n = (int)N, c = (int)C, hd = c / H, nq = (int)q_count;
    const float scale = 1.0f / sqrtf((float)hd);
    // x == NULL: qkv already holds the projections of all N tokens (e.g. all-gathered from the ranks that own them)
    if (x != nullptr) sgemm(x, C, 1, w_qkv, C, 1, b_qkv, qkv, 3 * C, n, 3 * c, c, st);
    if (attn_tc_supported(hd)) attn_tc_fwd(qkv, n, c, H, scale, o, lse, st, (int)q_start, (int)q_stride, nq);
    else switch (hd) {
        case 8: launch_fwd<8>(qkv, n, c, H, scale, o, lse, st, (int)q_start, (int)q_stride, nq); break;
        case 16: launch_fwd<16>(qkv, n, c, H, scale, o, lse, st, (int)q_start, (int)q_stride, nq); break;
        case 32: launch_fwd<32>(qkv, n, c, H, scale, o, lse, st, (int)q_start, (int)q_stride, nq); break;
        case 64: launch_fwd<64>(qkv, n, c, H, scale, o, lse, st, (int)q_start, (int)q_stride, nq); break;
        default: launch_fwd<128>(qkv, n, c, H, scale, o, lse, st, (int)q_start, (int)q_stride, nq); break;
    }
    sgemm(o, C, 1, w_proj, C, 1, b_proj, y, C, nq, c, c, st);
    MOMA_CUDA_LAUNCH_CHECK("attn_fwd_rows");
    note_launches(x != nullptr ? 3 : 2);
    return MOMA_OK;
}

// workspace: dO [N, C] | dqkv [N, 3C] | delta [H, N]
extern "C" __attribute__((visibility("default"))) size_t moma_attn_bwd_workspace_bytes(int64_t N, int64_t C, int H) {
    if (N <= 0 || C <= 0 || H <= 0) return 0;
    return (size_t)(N * C + N * 3 * C + (int64_t)H * N + 64) * sizeof(float);
}

extern "C" __attribute__((visibility("default"))) int moma_attn_bwd(const float* x, const float* w_qkv, const float* w_proj,
                             const float* qkv, const float* o, const float* lse,
                             const float* grad_y, int64_t N, int64_t C, int H, float* grad_x,
                             float* grad_w_qkv, float* grad_b_qkv, float* grad_w_proj,
                             float* grad_b_proj, void* workspace, size_t workspace_bytes,
                             moma_stream_t stream) {
    int rc = check_attn("attn_bwd", N, C, H);
    if (rc != MOMA_OK) return rc;
    MOMA_REQUIRE(x && w_qkv && w_proj && qkv && o && lse && grad_y, MOMA_ERR_INVALID, "attn_bwd: null pointer");
    MOMA_REQUIRE(workspace && workspace_bytes >= moma_attn_bwd_workspace_bytes(N, C, H), MOMA_ERR_WORKSPACE,
                 "attn_bwd: workspace too small");
    MOMA_REQUIRE(aligned16(workspace) && aligned16(qkv), MOMA_ERR_ALIGN, "attn_bwd: unaligned workspace/qkv");
    cudaStream_t st = as_stream(stream);
    const int n = (int)N, c = (int)C, hd = c / H;
    const float scale = 1.0f / sqrtf((float)hd);
    float* dO = static_cast<float*>(workspace);
    float* dqkv = dO + N * C;
    float* delta = dqkv + N * 3 * C;
    // The backward is a small DAG, not a chain: fork it over two side streams (captured as parallel branches
    // by a CUDA graph) so only  dO -> delta -> dQ || dK,dV -> dx  stays on the critical path.
    //   s1: dW_proj, db_proj          (need only dy, o)
    //   st: dO, delta, dQ, dx         s2: dK/dV, then dW_qkv, db_qkv (after dQ)
    BwdStreams& bs = bwd_streams();
    cudaStream_t s1 = bs.ok ? bs.s1 : st, s2 = bs.ok ? bs.s2 : st;
    if (bs.ok) { cudaEventRecord(bs.fork, st); cudaStreamWaitEvent(s1, bs.fork, 0); }
    const bool tc = attn_tc_supported(hd);
    bool split = false;
    if (!tc) switch (hd) {
        case 8: split = dkv_wants_split<8>(n, H); break;
        case 16: split = dkv_wants_split<16>(n, H); break;
        case 32: split = dkv_wants_split<32>(n, H); break;
        case 64: split = dkv_wants_split<64>(n, H); break;
        default: split = dkv_wants_split<128>(n, H); break;
    }
    if (split) {          // dK/dV are accumulated by two CTAs each: zero them early, off the critical path
        if (bs.ok) cudaStreamWaitEvent(s2, bs.fork, 0);
        cudaMemset2DAsync(dqkv + C, (size_t)3 * C * sizeof(float), 0, (size_t)2 * C * sizeof(float), (size_t)N, s2);
    }
    // proj backward: dW_proj[co, ci] = sum_n dy[n, co] o[n, ci];  db = colsum(dy);  dO = dy W_proj
    if (grad_b_proj) colsum_masked(grad_y, nullptr, n, c, grad_b_proj, s1);
    if (grad_w_proj) sgemm(grad_y, 1, C, o, 1, C, nullptr, grad_w_proj, C, c, c, n, s1);
    sgemm(grad_y, C, 1, w_proj, 1, C, nullptr, dO, C, n, c, c, st);
    launch_pdl(attn_delta_kernel, dim3((n * H + 3) / 4), dim3(128), 0, st, dO, o, n, c, H, delta);
    if (bs.ok) { cudaEventRecord(bs.d_o, st); cudaStreamWaitEvent(s2, bs.d_o, 0); }
    if (tc) attn_tc_bwd(qkv, dO, lse, delta, n, c, H, scale, dqkv, st, s2);
    else switch (hd) {
        case 8: launch_bwd<8>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st, s2, split); break;
        case 16: launch_bwd<16>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st, s2, split); break;
        case 32: launch_bwd<32>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st, s2, split); break;
        case 64: launch_bwd<64>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st, s2, split); break;
        default: launch_bwd<128>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st, s2, split); break;
    }
    if (bs.ok) {
        cudaEventRecord(bs.dq, st);  cudaStreamWaitEvent(s2, bs.dq, 0);     // s2 now has dQ and dK/dV
        cudaEventRecord(bs.dkv, s2); cudaStreamWaitEvent(st, bs.dkv, 0);    // st too
    }
    // qkv backward: dW_qkv[j, ci] = sum_n dqkv[n, j] x[n, ci]; db = colsum(dqkv); dx = dqkv W_qkv
    if (grad_b_qkv) colsum_masked(dqkv, nullptr, n, 3 * c, grad_b_qkv, s2);
    if (grad_w_qkv) sgemm(dqkv, 1, 3 * C, x, 1, C, nullptr, grad_w_qkv, C, 3 * c, c, n, s2);
    if (grad_x) sgemm(dqkv, 3 * C, 1, w_qkv, 1, C, nullptr, grad_x, C, n, c, 3 * c, st);
    if (bs.ok) {                                                             // join
        cudaEventRecord(bs.join1, s1); cudaStreamWaitEvent(st, bs.join1, 0);
        cudaEventRecord(bs.join2, s2); cudaStreamWaitEvent(st, bs.join2, 0);
    }
    MOMA_CUDA_LAUNCH_CHECK("attn_bwd");
    note_launches(4 + (grad_w_proj != nullptr) + (grad_b_proj != nullptr) + (grad_w_qkv != nullptr) +
                  (grad_b_qkv != nullptr) + (grad_x != nullptr));
    return MOMA_OK;
}
