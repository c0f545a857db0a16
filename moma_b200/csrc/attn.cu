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
#include "common.cuh"

namespace moma {

// ------------------------------------------------------------------ generic SGEMM
// C[m, n] = sum_k A(m,k) * B(n,k) (+ bias[n]);  A(m,k) = A[m*a_rs + k*a_cs], same for B.
// 64x64x16 tiles, 256 threads, 4x4 micro-tile per thread.
template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int64_t a_rs, int64_t a_cs, const float* __restrict__ Bm,
             int64_t b_rs, int64_t b_cs, const float* __restrict__ bias, float* __restrict__ Cm,
             int64_t ldc, int M, int N, int K) {
    __shared__ float As[16][64 + 1];
    __shared__ float Bs[16][64 + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int i = tid + it * 256;
            {
                const int r = A_KCONTIG ? (i >> 4) : (i & 63);
                const int k = A_KCONTIG ? (i & 15) : (i >> 6);
                As[k][r] = (m0 + r < M && k0 + k < K) ? A[(int64_t)(m0 + r) * a_rs + (int64_t)(k0 + k) * a_cs] : 0.f;
            }
            {
                const int r = B_KCONTIG ? (i >> 4) : (i & 63);
                const int k = B_KCONTIG ? (i & 15) : (i >> 6);
                Bs[k][r] = (n0 + r < N && k0 + k < K) ? Bm[(int64_t)(n0 + r) * b_rs + (int64_t)(k0 + k) * b_cs] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty + 16 * i]; b[i] = Bs[k][tx + 16 * i]; }
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
        if (r >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx + 16 * j;
            if (c < N) Cm[(int64_t)r * ldc + c] = acc[i][j] + (bias ? bias[c] : 0.f);
        }
    }
}

static void sgemm(const float* A, int64_t a_rs, int64_t a_cs, const float* Bm, int64_t b_rs,
                  int64_t b_cs, const float* bias, float* Cm, int64_t ldc, int M, int N, int K,
                  cudaStream_t st) {
    const dim3 grid((N + 63) / 64, (M + 63) / 64);
    const bool ak = (a_cs == 1), bk = (b_cs == 1);
    if (ak && bk) sgemm_kernel<true, true><<<grid, 256, 0, st>>>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K);
    else if (ak && !bk) sgemm_kernel<true, false><<<grid, 256, 0, st>>>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K);
    else if (!ak && bk) sgemm_kernel<false, true><<<grid, 256, 0, st>>>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K);
    else sgemm_kernel<false, false><<<grid, 256, 0, st>>>(A, a_rs, a_cs, Bm, b_rs, b_cs, bias, Cm, ldc, M, N, K);
}

// column sums: out[c] = sum_r X[r, c]   (bias gradients)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int rows, int cols, float* __restrict__ out) {
    __shared__ float red[8][32 + 1];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    float s = 0.f;
    if (c < cols)
        for (int r = ry; r < rows; r += 8) s += X[(int64_t)r * cols + c];
    red[ry][threadIdx.x & 31] = s;
    __syncthreads();
    if (ry == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        out[c] = t;
    }
}

// ------------------------------------------------------------------ fused attention core
// Tiles: 32 "row" tokens per CTA (8 per warp, 4 warps), 64 "column" tokens per step.
constexpr int kAR = 32;    // rows per CTA
constexpr int kAC = 64;    // columns per inner tile
constexpr int kAThreads = 128;
constexpr int kLdP = kAC + 4;

template <int HD> struct AttnCfg {
    static constexpr int LD = HD + 4;                       // padded smem row (floats)
    static constexpr int CPL = HD >= 32 ? HD / 32 : 1;      // output columns per lane
    static constexpr int RPL = HD >= 32 ? 8 : HD / 4;       // output rows per lane (32/HD row groups)
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

// s[r][cc] = sum_d a[warp*8 + r][d] * b[lane + 32*cc][d]
template <int HD>
__device__ __forceinline__ void dot_tile(const float* a_s, const float* b_s, int warp, int lane,
                                         float (&s)[8][2]) {
    constexpr int LD = HD + 4;
#pragma unroll
    for (int r = 0; r < 8; ++r) { s[r][0] = 0.f; s[r][1] = 0.f; }
#pragma unroll 4
    for (int v = 0; v < HD / 4; ++v) {
        const float4 b0 = *reinterpret_cast<const float4*>(b_s + lane * LD + 4 * v);
        const float4 b1 = *reinterpret_cast<const float4*>(b_s + (lane + 32) * LD + 4 * v);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(a_s + (warp * 8 + r) * LD + 4 * v);
            s[r][0] += a.x * b0.x + a.y * b0.y + a.z * b0.z + a.w * b0.w;
            s[r][1] += a.x * b1.x + a.y * b1.y + a.z * b1.z + a.w * b1.w;
        }
    }
}

// acc[r][c] += sum_j p[row(r)][j] * v[j][col(c)],  j over the 64-column tile
template <int HD>
__device__ __forceinline__ void acc_tile(const float* p_s, const float* v_s, int warp, int lane,
                                         float (&acc)[AttnCfg<HD>::RPL][AttnCfg<HD>::CPL]) {
    using Cfg = AttnCfg<HD>;
    constexpr int LD = Cfg::LD;
    const int rbase = warp * 8 + Cfg::roff(lane);
    const int cbase = Cfg::cbase(lane);
#pragma unroll 2
    for (int j = 0; j < kAC; j += 4) {
        float vv[4][Cfg::CPL];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c) vv[jj][c] = v_s[(j + jj) * LD + cbase + 32 * c];
#pragma unroll
        for (int r = 0; r < Cfg::RPL; ++r) {
            const float4 p = *reinterpret_cast<const float4*>(p_s + (rbase + r) * kLdP + j);
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c)
                acc[r][c] += p.x * vv[0][c] + p.y * vv[1][c] + p.z * vv[2][c] + p.w * vv[3][c];
        }
    }
}

// ---- forward: grid (ceil(N/32), H)
template <int HD>
__global__ void __launch_bounds__(kAThreads)
attn_fwd_kernel(const float* __restrict__ qkv, int N, int C, float scale, float* __restrict__ o,
                float* __restrict__ lse) {
    using Cfg = AttnCfg<HD>;
    extern __shared__ __align__(16) float sm[];
    float* q_s = sm;                         // [32][LD]
    float* k_s = q_s + kAR * Cfg::LD;        // [64][LD]
    float* v_s = k_s + kAC * Cfg::LD;        // [64][LD]
    float* p_s = v_s + kAC * Cfg::LD;        // [32][kLdP]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.y, i0 = blockIdx.x * kAR;
    const int64_t ldg = 3 * (int64_t)C;
    const float* qg = qkv + h * HD;
    const float* kg = qkv + C + h * HD;
    const float* vg = qkv + 2 * C + h * HD;

    load_rows<HD>(q_s, qg, ldg, i0, kAR, N);
    float m_run[8], l_run[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) { m_run[r] = -CUDART_INF_F; l_run[r] = 0.f; }
    float acc[Cfg::RPL][Cfg::CPL] = {};

    for (int j0 = 0; j0 < N; j0 += kAC) {
        __syncthreads();
        load_rows<HD>(k_s, kg, ldg, j0, kAC, N);
        load_rows<HD>(v_s, vg, ldg, j0, kAC, N);
        __syncthreads();
        float s[8][2];
        dot_tile<HD>(q_s, k_s, warp, lane, s);
        float corr[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float s0 = (j0 + lane < N) ? s[r][0] * scale : -CUDART_INF_F;
            const float s1 = (j0 + lane + 32 < N) ? s[r][1] * scale : -CUDART_INF_F;
            const float mx = warp_max(fmaxf(s0, s1));
            const float m_new = fmaxf(m_run[r], mx);
            corr[r] = (m_run[r] == -CUDART_INF_F) ? 0.f : expf(m_run[r] - m_new);
            const float p0 = (s0 == -CUDART_INF_F) ? 0.f : expf(s0 - m_new);
            const float p1 = (s1 == -CUDART_INF_F) ? 0.f : expf(s1 - m_new);
            p_s[(warp * 8 + r) * kLdP + lane] = p0;
            p_s[(warp * 8 + r) * kLdP + lane + 32] = p1;
            l_run[r] = l_run[r] * corr[r] + warp_sum(p0 + p1);
            m_run[r] = m_new;
        }
        const int roff = Cfg::roff(lane);
#pragma unroll
        for (int r = 0; r < Cfg::RPL; ++r) {
            // corr is warp-uniform per row; pick this lane's rows
            float cr = corr[0];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) cr = (rr == r + roff) ? corr[rr] : cr;
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c) acc[r][c] *= cr;
        }
        __syncwarp();
        acc_tile<HD>(p_s, v_s, warp, lane, acc);
    }
    const int roff = Cfg::roff(lane);
    const int cbase = Cfg::cbase(lane);
#pragma unroll
    for (int r = 0; r < Cfg::RPL; ++r) {
        float lr = l_run[0];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) lr = (rr == r + roff) ? l_run[rr] : lr;
        const int row = i0 + warp * 8 + roff + r;
        if (row < N) {
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c)
                o[(int64_t)row * C + h * HD + cbase + 32 * c] = acc[r][c] / lr;
        }
    }
    if (lane < 8) {
        float mr = m_run[0], lr = l_run[0];
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) { mr = (rr == lane) ? m_run[rr] : mr; lr = (rr == lane) ? l_run[rr] : lr; }
        const int row = i0 + warp * 8 + lane;
        if (row < N) lse[(int64_t)h * N + row] = mr + logf(lr);
    }
}

// attention probabilities for Attention_viz: probs[h, i, j] = exp(s_ij - lse_i)
__global__ void __launch_bounds__(256)
attn_probs_kernel(const float* __restrict__ qkv, const float* __restrict__ lse, int N, int C, int H,
                  float scale, float* __restrict__ probs) {
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
template <int HD>
__global__ void __launch_bounds__(kAThreads)
attn_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ dO,
                   const float* __restrict__ lse, const float* __restrict__ delta, int N, int C,
                   float scale, float* __restrict__ dqkv) {
    using Cfg = AttnCfg<HD>;
    extern __shared__ __align__(16) float sm[];
    float* q_s = sm;                          // [32][LD]
    float* do_s = q_s + kAR * Cfg::LD;        // [32][LD]
    float* k_s = do_s + kAR * Cfg::LD;        // [64][LD]
    float* v_s = k_s + kAC * Cfg::LD;         // [64][LD]
    float* p_s = v_s + kAC * Cfg::LD;         // [32][kLdP]  (holds dS)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.y, i0 = blockIdx.x * kAR;
    const int64_t ldg = 3 * (int64_t)C;
    load_rows<HD>(q_s, qkv + h * HD, ldg, i0, kAR, N);
    load_rows<HD>(do_s, dO + h * HD, C, i0, kAR, N);
    float lse_r[8], del_r[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = i0 + warp * 8 + r;
        lse_r[r] = row < N ? lse[(int64_t)h * N + row] : 0.f;
        del_r[r] = row < N ? delta[(int64_t)h * N + row] : 0.f;
    }
    float acc[Cfg::RPL][Cfg::CPL] = {};
    for (int j0 = 0; j0 < N; j0 += kAC) {
        __syncthreads();
        load_rows<HD>(k_s, qkv + C + h * HD, ldg, j0, kAC, N);
        load_rows<HD>(v_s, qkv + 2 * C + h * HD, ldg, j0, kAC, N);
        __syncthreads();
        float s[8][2], dp[8][2];
        dot_tile<HD>(q_s, k_s, warp, lane, s);
        dot_tile<HD>(do_s, v_s, warp, lane, dp);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const bool ok = (j0 + lane + 32 * cc < N);
                const float p = ok ? expf(s[r][cc] * scale - lse_r[r]) : 0.f;
                p_s[(warp * 8 + r) * kLdP + lane + 32 * cc] = p * (dp[r][cc] - del_r[r]);
            }
        }
        __syncwarp();
        acc_tile<HD>(p_s, k_s, warp, lane, acc);
    }
    const int roff = Cfg::roff(lane);
    const int cbase = Cfg::cbase(lane);
#pragma unroll
    for (int r = 0; r < Cfg::RPL; ++r) {
        const int row = i0 + warp * 8 + roff + r;
        if (row < N)
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c)
                dqkv[(int64_t)row * ldg + h * HD + cbase + 32 * c] = acc[r][c] * scale;
    }
}

// dK, dV: grid (ceil(N/32) key blocks, H); rows = keys, columns = queries
template <int HD>
__global__ void __launch_bounds__(kAThreads)
attn_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ dO,
                    const float* __restrict__ lse, const float* __restrict__ delta, int N, int C,
                    float scale, float* __restrict__ dqkv) {
    using Cfg = AttnCfg<HD>;
    extern __shared__ __align__(16) float sm[];
    float* k_s = sm;                           // [32][LD]  own keys
    float* v_s = k_s + kAR * Cfg::LD;          // [32][LD]  own values
    float* q_s = v_s + kAR * Cfg::LD;          // [64][LD]  query tile
    float* do_s = q_s + kAC * Cfg::LD;         // [64][LD]  dO tile
    float* p_s = do_s + kAC * Cfg::LD;         // [32][kLdP]  P^T
    float* ds_s = p_s + kAR * kLdP;            // [32][kLdP]  dS^T
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int h = blockIdx.y, j0 = blockIdx.x * kAR;
    const int64_t ldg = 3 * (int64_t)C;
    load_rows<HD>(k_s, qkv + C + h * HD, ldg, j0, kAR, N);
    load_rows<HD>(v_s, qkv + 2 * C + h * HD, ldg, j0, kAR, N);
    float acc_k[Cfg::RPL][Cfg::CPL] = {};
    float acc_v[Cfg::RPL][Cfg::CPL] = {};
    for (int i0 = 0; i0 < N; i0 += kAC) {
        __syncthreads();
        load_rows<HD>(q_s, qkv + h * HD, ldg, i0, kAC, N);
        load_rows<HD>(do_s, dO + h * HD, C, i0, kAC, N);
        __syncthreads();
        float st[8][2], dpt[8][2];
        dot_tile<HD>(k_s, q_s, warp, lane, st);      // st[r][cc] = k_j . q_i
        dot_tile<HD>(v_s, do_s, warp, lane, dpt);    // dpt      = v_j . do_i
        float lse_c[2], del_c[2];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int i = i0 + lane + 32 * cc;
            lse_c[cc] = i < N ? lse[(int64_t)h * N + i] : 0.f;
            del_c[cc] = i < N ? delta[(int64_t)h * N + i] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const bool rok = (j0 + warp * 8 + r < N);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const bool ok = rok && (i0 + lane + 32 * cc < N);
                const float p = ok ? expf(st[r][cc] * scale - lse_c[cc]) : 0.f;
                p_s[(warp * 8 + r) * kLdP + lane + 32 * cc] = p;
                ds_s[(warp * 8 + r) * kLdP + lane + 32 * cc] = p * (dpt[r][cc] - del_c[cc]);
            }
        }
        __syncwarp();
        acc_tile<HD>(p_s, do_s, warp, lane, acc_v);
        acc_tile<HD>(ds_s, q_s, warp, lane, acc_k);
    }
    const int roff = Cfg::roff(lane);
    const int cbase = Cfg::cbase(lane);
#pragma unroll
    for (int r = 0; r < Cfg::RPL; ++r) {
        const int row = j0 + warp * 8 + roff + r;
        if (row < N)
#pragma unroll
            for (int c = 0; c < Cfg::CPL; ++c) {
                dqkv[(int64_t)row * ldg + C + h * HD + cbase + 32 * c] = acc_k[r][c] * scale;
                dqkv[(int64_t)row * ldg + 2 * C + h * HD + cbase + 32 * c] = acc_v[r][c];
            }
    }
}

template <int HD> static size_t fwd_smem() { return (size_t)((kAR + 2 * kAC) * (HD + 4) + kAR * kLdP) * sizeof(float); }
template <int HD> static size_t dq_smem() { return (size_t)((2 * kAR + 2 * kAC) * (HD + 4) + kAR * kLdP) * sizeof(float); }
template <int HD> static size_t dkv_smem() { return (size_t)((2 * kAR + 2 * kAC) * (HD + 4) + 2 * kAR * kLdP) * sizeof(float); }

template <int HD>
static void launch_fwd(const float* qkv, int N, int C, int H, float scale, float* o, float* lse, cudaStream_t st) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<HD>()); attr = true; }
    attn_fwd_kernel<HD><<<dim3((N + kAR - 1) / kAR, H), kAThreads, fwd_smem<HD>(), st>>>(qkv, N, C, scale, o, lse);
}
template <int HD>
static void launch_bwd(const float* qkv, const float* dO, const float* lse, const float* delta, int N, int C, int H,
                       float scale, float* dqkv, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dq_smem<HD>());
        cudaFuncSetAttribute(attn_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dkv_smem<HD>());
        attr = true;
    }
    const dim3 grid((N + kAR - 1) / kAR, H);
    attn_bwd_dq_kernel<HD><<<grid, kAThreads, dq_smem<HD>(), st>>>(qkv, dO, lse, delta, N, C, scale, dqkv);
    attn_bwd_dkv_kernel<HD><<<grid, kAThreads, dkv_smem<HD>(), st>>>(qkv, dO, lse, delta, N, C, scale, dqkv);
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

extern "C" __attribute__((visibility("default"))) int moma_attn_fwd(const float* x, const float* w_qkv, const float* b_qkv,
                             const float* w_proj, const float* b_proj, int64_t N, int64_t C, int H,
                             float* y, float* qkv, float* o, float* lse, float* attn_probs,
                             moma_stream_t stream) {
    int rc = check_attn("attn_fwd", N, C, H);
    if (rc != MOMA_OK) return rc;
    MOMA_REQUIRE(x && w_qkv && w_proj && b_proj && y && qkv && o && lse, MOMA_ERR_INVALID, "attn_fwd: null pointer");
    MOMA_REQUIRE(aligned16(qkv) && aligned16(o), MOMA_ERR_ALIGN, "attn_fwd: qkv/o must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const int n = (int)N, c = (int)C, hd = c / H;
    const float scale = 1.0f / sqrtf((float)hd);
    sgemm(x, C, 1, w_qkv, C, 1, b_qkv, qkv, 3 * C, n, 3 * c, c, st);
    switch (hd) {
        case 8: launch_fwd<8>(qkv, n, c, H, scale, o, lse, st); break;
        case 16: launch_fwd<16>(qkv, n, c, H, scale, o, lse, st); break;
        case 32: launch_fwd<32>(qkv, n, c, H, scale, o, lse, st); break;
        case 64: launch_fwd<64>(qkv, n, c, H, scale, o, lse, st); break;
        default: launch_fwd<128>(qkv, n, c, H, scale, o, lse, st); break;
    }
    if (attn_probs) {
        const int64_t tot = (int64_t)H * N * N;
        attn_probs_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(qkv, lse, n, c, H, scale, attn_probs);
    }
    sgemm(o, C, 1, w_proj, C, 1, b_proj, y, C, n, c, c, st);
    MOMA_CUDA_LAUNCH_CHECK("attn_fwd");
    note_launches(attn_probs ? 4 : 3);
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
    // proj backward: dW_proj[co, ci] = sum_n dy[n, co] o[n, ci];  db = colsum(dy);  dO = dy W_proj
    if (grad_w_proj) sgemm(grad_y, 1, C, o, 1, C, nullptr, grad_w_proj, C, c, c, n, st);
    if (grad_b_proj) colsum_kernel<<<(c + 31) / 32, 256, 0, st>>>(grad_y, n, c, grad_b_proj);
    sgemm(grad_y, C, 1, w_proj, 1, C, nullptr, dO, C, n, c, c, st);
    attn_delta_kernel<<<(n * H + 3) / 4, 128, 0, st>>>(dO, o, n, c, H, delta);
    switch (hd) {
        case 8: launch_bwd<8>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st); break;
        case 16: launch_bwd<16>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st); break;
        case 32: launch_bwd<32>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st); break;
        case 64: launch_bwd<64>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st); break;
        default: launch_bwd<128>(qkv, dO, lse, delta, n, c, H, scale, dqkv, st); break;
    }
    // qkv backward: dW_qkv[j, ci] = sum_n dqkv[n, j] x[n, ci]; db = colsum(dqkv); dx = dqkv W_qkv
    if (grad_w_qkv) sgemm(dqkv, 1, 3 * C, x, 1, C, nullptr, grad_w_qkv, C, 3 * c, c, n, st);
    if (grad_b_qkv) colsum_kernel<<<(3 * c + 31) / 32, 256, 0, st>>>(dqkv, n, 3 * c, grad_b_qkv);
    if (grad_x) sgemm(dqkv, 3 * C, 1, w_qkv, 1, C, nullptr, grad_x, C, n, c, 3 * c, st);
    MOMA_CUDA_LAUNCH_CHECK("attn_bwd");
    note_launches(4 + (grad_w_proj != nullptr) + (grad_b_proj != nullptr) + (grad_w_qkv != nullptr) +
                  (grad_b_qkv != nullptr) + (grad_x != nullptr));
    return MOMA_OK;
}
