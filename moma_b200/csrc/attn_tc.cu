// Attention core on the tensor cores: scores, softmax and value product of MoMA/criterion_moco_att.py:158-164 and their
// backward, as flash-style kernels whose contractions run as 3xTF32 warp-level MMAs (mma.sync.m16n8k8).
//
// Why warp-level MMA and 3xTF32: the attention of this path is tiny -- N = batch = 256..1024 tokens, head_dim 16..32, so
// 0.03..0.5 GFLOP per module -- and it has to hold the FP32 parity bar (1e-5) because its output is the query / key of
// the InfoNCE loss; a 128-row tcgen05 tile has nothing to chew on at head_dim 16, while one warp per 16 query rows gives
// 256 independent warps at N = 512, H = 8.  x = hi + lo (hi = TF32 rounding, lo = exact remainder), a.b ~= a_lo.b_hi +
// a_hi.b_lo + a_hi.b_hi reproduces fp32 products (csrc/gemm.cu has the error analysis); every per-tile product starts
// from a zero accumulator and is added to the running sum with an ordinary fp32 add, because the tensor core adds with
// truncation and a long accumulation chain would drift.
//
// Fragment trick: the score tile comes out of the first MMA in the accumulator layout (thread holds columns 2t, 2t+1 of an
// 8-key group) but the second MMA wants it in the A layout (columns t, t+4).  A contraction over keys does not care about
// their order, so the 8 keys of a group are simply consumed in the order (0,2,4,6,1,3,5,7): the accumulator registers ARE
// the A fragment, and the matching B fragment reads rows 2t and 2t+1 of the value tile.  No shuffles, no shared memory.
//
// Kernels (grid = (ceil(rows / 64), H), 128 threads = 4 warps x 16 rows, K/V or Q/dO tiles of 64 rows double-buffered in
// shared memory with cp.async):
//   attn_fwd_tc_kernel      O = softmax(Q K^T * scale) V, lse                      (also for a strided subset of the rows)
//   attn_bwd_dq_tc_kernel   dQ = [P o (dO V^T - delta)] K * scale
//   attn_bwd_dkv_tc_kernel  dV = P^T dO,  dK = [P o (dO V^T - delta)]^T Q * scale
// head_dim 8, 16, 32, 64 (all operand fragments of a warp live in registers; at 64 the dK/dV kernel spills ~0.7 KB per
// thread and still beats the CUDA-core kernel); head_dim 128 uses the SIMT kernels of attn.cu.
#include <math_constants.h>
#include <cstdlib>
#include <string>
#include "common.cuh"

namespace moma {
namespace atc {

constexpr int kRows = 64, kTile = 64, kThreads = 128;
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// acc += a . b with fp32-level accuracy (small terms first)
__device__ __forceinline__ void mma3(float (&acc)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    mma_tf32(acc, al, bh);
    mma_tf32(acc, ah, bl);
    mma_tf32(acc, ah, bh);
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// 64 rows x HD floats, global (row stride `rs` floats) -> shared [64][HD + 4]; rows >= n_valid are zero-filled
template <int HD, int THREADS = kThreads>
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, int64_t rs, int row0, int n_valid) {
    constexpr int LD = HD + 4, V = HD / 4;
#pragma unroll
    for (int i = threadIdx.x; i < kTile * V; i += THREADS) {
        const int r = i / V, v = i - r * V;
        const bool ok = row0 + r < n_valid;
        cp_async16(dst + r * LD + 4 * v, ok ? src + (int64_t)(row0 + r) * rs + 4 * v : src, ok);
    }
}
// B fragment of a contraction over the tile's COLUMNS: B(k, n) = tile[n0 + n][k0 + k]  (b0: k = t, b1: k = t + 4; n = g)
template <int LD>
__device__ __forceinline__ void frag_cols(const float* tile, int n0, int k0, int g, int t, uint32_t (&hi)[2], uint32_t (&lo)[2]) {
    const float* p = tile + (n0 + g) * LD + k0 + t;
    split_tf32(p[0], hi[0], lo[0]);
    split_tf32(p[4], hi[1], lo[1]);
}
// B fragment of a contraction over the tile's ROWS in the permuted order (see the header): B(k, n) = tile[r0 + pi(k)][n0 + n]
template <int LD>
__device__ __forceinline__ void frag_rows(const float* tile, int r0, int n0, int g, int t, uint32_t (&hi)[2], uint32_t (&lo)[2]) {
    const float* p = tile + (r0 + 2 * t) * LD + n0 + g;
    split_tf32(p[0], hi[0], lo[0]);
    split_tf32(p[LD], hi[1], lo[1]);
}
// accumulator layout (c0 c1 | c2 c3) of an 8-column group -> A fragment (a0 a1 a2 a3) under the permuted column order
__device__ __forceinline__ void acc_to_a(const float (&c)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
    split_tf32(c[0], hi[0], lo[0]);
    split_tf32(c[2], hi[1], lo[1]);
    split_tf32(c[1], hi[2], lo[2]);
    split_tf32(c[3], hi[3], lo[3]);
}
// A fragments (all k-steps) of 16 rows read straight from global memory: row r_lo / r_hi, HD contiguous floats, times `mul`
template <int HD>
__device__ __forceinline__ void load_a_rows(const float* __restrict__ p_lo, const float* __restrict__ p_hi, bool ok_lo, bool ok_hi,
                                            float mul, int t, uint32_t (&hi)[HD / 8][4], uint32_t (&lo)[HD / 8][4]) {
#pragma unroll
    for (int ks = 0; ks < HD / 8; ++ks) {
        const float a0 = ok_lo ? p_lo[ks * 8 + t] * mul : 0.f, a1 = ok_hi ? p_hi[ks * 8 + t] * mul : 0.f;
        const float a2 = ok_lo ? p_lo[ks * 8 + t + 4] * mul : 0.f, a3 = ok_hi ? p_hi[ks * 8 + t + 4] * mul : 0.f;
        split_tf32(a0, hi[ks][0], lo[ks][0]);
        split_tf32(a1, hi[ks][1], lo[ks][1]);
        split_tf32(a2, hi[ks][2], lo[ks][2]);
        split_tf32(a3, hi[ks][3], lo[ks][3]);
    }
}

// ---------------------------------------------------------------------------------------------------- forward
// WARPS: 16 query rows each; 4 for the usual grids, 2 or 1 when rows x heads would leave most SMs idle
template <int HD, int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
attn_fwd_tc_kernel(const float* __restrict__ qkv, int N, int C, float scale, float* __restrict__ o, float* __restrict__ lse,
                   int q_start, int q_stride, int NQ) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int LD = HD + 4, KS = HD / 8, NT = HD / 8, THREADS = 32 * WARPS;
    extern __shared__ __align__(16) float smem[];
    constexpr int STAGE = 2 * kTile * LD;                       // K tile then V tile
    const int h = blockIdx.y, q0 = blockIdx.x * (16 * WARPS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;
    const int64_t ld3 = 3ll * C;
    const float* kbase = qkv + C + h * HD;
    const float* vbase = qkv + 2 * C + h * HD;

    uint32_t qh[KS][4], ql[KS][4];
    load_a_rows<HD>(qkv + (int64_t)(q_start + (int64_t)r_lo * q_stride) * ld3 + h * HD,
                    qkv + (int64_t)(q_start + (int64_t)r_hi * q_stride) * ld3 + h * HD, r_lo < NQ, r_hi < NQ,
                    scale * kLog2e, t, qh, ql);
    float oacc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) { oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f; }
    float m_lo = -CUDART_INF_F, m_hi = -CUDART_INF_F, l_lo = 0.f, l_hi = 0.f;

    const int ntiles = (N + kTile - 1) / kTile;
    load_tile<HD, THREADS>(smem, kbase, ld3, 0, N);
    load_tile<HD, THREADS>(smem + kTile * LD, vbase, ld3, 0, N);
    cp_async_commit();
    for (int kt = 0; kt < ntiles; ++kt) {
        cp_async_wait<0>();
        __syncthreads();                               // tile kt visible; everyone is done with tile kt - 1
        if (kt + 1 < ntiles) {
            float* nxt = smem + ((kt + 1) & 1) * STAGE;
            load_tile<HD, THREADS>(nxt, kbase, ld3, (kt + 1) * kTile, N);
            load_tile<HD, THREADS>(nxt + kTile * LD, vbase, ld3, (kt + 1) * kTile, N);
            cp_async_commit();
        }
        const float* K = smem + (kt & 1) * STAGE;
        const float* V = K + kTile * LD;
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t bh[2], bl[2];
                frag_cols<LD>(K, j * 8, ks * 8, g, t, bh, bl);
                mma3(s[j], qh[ks], ql[ks], bh, bl);
            }
        if (kt == ntiles - 1 && (N % kTile) != 0) {     // keys beyond N
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = kt * kTile + j * 8 + 2 * t;
                if (c >= N) { s[j][0] = -CUDART_INF_F; s[j][2] = -CUDART_INF_F; }
                if (c + 1 >= N) { s[j][1] = -CUDART_INF_F; s[j][3] = -CUDART_INF_F; }
            }
        }
        float mx_lo = -CUDART_INF_F, mx_hi = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx_lo = fmaxf(mx_lo, fmaxf(s[j][0], s[j][1]));
            mx_hi = fmaxf(mx_hi, fmaxf(s[j][2], s[j][3]));
        }
        mx_lo = quad_max(mx_lo); mx_hi = quad_max(mx_hi);
        const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);       // finite: every tile has a valid key
        const float c_lo = ex2f(m_lo - mn_lo), c_hi = ex2f(m_hi - mn_hi);        // exp2(-inf) = 0 on the first tile
        m_lo = mn_lo; m_hi = mn_hi;
        float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = ex2f(s[j][0] - mn_lo); s[j][1] = ex2f(s[j][1] - mn_lo);
            s[j][2] = ex2f(s[j][2] - mn_hi); s[j][3] = ex2f(s[j][3] - mn_hi);
            sum_lo += s[j][0] + s[j][1]; sum_hi += s[j][2] + s[j][3];
        }
        l_lo = l_lo * c_lo + sum_lo; l_hi = l_hi * c_hi + sum_hi;                 // per-thread partial sums (quad-reduced at the end)
        float ot[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n) { ot[n][0] = ot[n][1] = ot[n][2] = ot[n][3] = 0.f; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t ph[4], pl[4];
            acc_to_a(s[j], ph, pl);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                uint32_t bh[2], bl[2];
                frag_rows<LD>(V, j * 8, n * 8, g, t, bh, bl);
                mma3(ot[n], ph, pl, bh, bl);
            }
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            oacc[n][0] = oacc[n][0] * c_lo + ot[n][0]; oacc[n][1] = oacc[n][1] * c_lo + ot[n][1];
            oacc[n][2] = oacc[n][2] * c_hi + ot[n][2]; oacc[n][3] = oacc[n][3] * c_hi + ot[n][3];
        }
    }
    l_lo = quad_sum(l_lo); l_hi = quad_sum(l_hi);
    const float i_lo = 1.0f / l_lo, i_hi = 1.0f / l_hi;
    if (r_lo < NQ) {
#pragma unroll
        for (int n = 0; n < NT; ++n)
            *reinterpret_cast<float2*>(o + (int64_t)r_lo * C + h * HD + n * 8 + 2 * t) = make_float2(oacc[n][0] * i_lo, oacc[n][1] * i_lo);
        if (t == 0) lse[(int64_t)h * NQ + r_lo] = (m_lo + log2f(l_lo)) * kLn2;
    }
    if (r_hi < NQ) {
#pragma unroll
        for (int n = 0; n < NT; ++n)
            *reinterpret_cast<float2*>(o + (int64_t)r_hi * C + h * HD + n * 8 + 2 * t) = make_float2(oacc[n][2] * i_hi, oacc[n][3] * i_hi);
        if (t == 0) lse[(int64_t)h * NQ + r_hi] = (m_hi + log2f(l_hi)) * kLn2;
    }
}

// ---------------------------------------------------------------------------------------------------- forward, keys split over the warps
// 16 query rows per CTA; its four warps walk DISJOINT quarters of the key / value tiles (warp w: tiles w, w + 4, ...), each
// with its own double-buffered tiles in shared memory and no CTA-wide barrier in the loop, and merge their (max, sum, O)
// through shared memory at the end: the serial chain per warp is a quarter of attn_fwd_tc_kernel's and the grid is 4x larger.
// Used from four key tiles up -- most valuable for few query rows against many keys, the K-sharded queue's owned-rows
// attention (512 rows x (W x 512) keys).
template <int HD>
__global__ void __launch_bounds__(128)
attn_fwd_splitkv_kernel(const float* __restrict__ qkv, int N, int C, float scale, float* __restrict__ o, float* __restrict__ lse,
                        int q_start, int q_stride, int NQ) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int LD = HD + 4, KS = HD / 8, NT = HD / 8, V4 = HD / 4;
    extern __shared__ __align__(16) float smem[];
    constexpr int STAGE = 2 * kTile * LD;                       // K tile then V tile
    const int h = blockIdx.y, q0 = blockIdx.x * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r_lo = q0 + g, r_hi = r_lo + 8;
    const int64_t ld3 = 3ll * C;
    const float* kbase = qkv + C + h * HD;
    const float* vbase = qkv + 2 * C + h * HD;
    float* wsm = smem + warp * 2 * STAGE;                       // this warp's two stages

    uint32_t qh[KS][4], ql[KS][4];
    load_a_rows<HD>(qkv + (int64_t)(q_start + (int64_t)r_lo * q_stride) * ld3 + h * HD,
                    qkv + (int64_t)(q_start + (int64_t)r_hi * q_stride) * ld3 + h * HD, r_lo < NQ, r_hi < NQ,
                    scale * kLog2e, t, qh, ql);
    float oacc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) { oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f; }
    float m_lo = -CUDART_INF_F, m_hi = -CUDART_INF_F, l_lo = 0.f, l_hi = 0.f;

    const int ntiles = (N + kTile - 1) / kTile;
    auto load_warp_tile = [&](float* dst, int kt) {             // K and V rows [kt * 64, +64) by this warp's 32 lanes
#pragma unroll
        for (int i = lane; i < kTile * V4; i += 32) {
            const int r = i / V4, v = i - r * V4;
            const bool ok = kt * kTile + r < N;
            const int64_t off = (int64_t)(kt * kTile + r) * ld3 + 4 * v;
            cp_async16(dst + r * LD + 4 * v, ok ? kbase + off : kbase, ok);
            cp_async16(dst + kTile * LD + r * LD + 4 * v, ok ? vbase + off : vbase, ok);
        }
    };
    if (warp < ntiles) load_warp_tile(wsm, warp);
    cp_async_commit();
    int it = 0;
    for (int kt = warp; kt < ntiles; kt += 4, ++it) {
        cp_async_wait<0>();
        __syncwarp();                                  // tile kt visible to the warp; every lane is done with the previous one
        if (kt + 4 < ntiles) load_warp_tile(wsm + ((it + 1) & 1) * STAGE, kt + 4);
        cp_async_commit();
        const float* K = wsm + (it & 1) * STAGE;
        const float* V = K + kTile * LD;
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t bh[2], bl[2];
                frag_cols<LD>(K, j * 8, ks * 8, g, t, bh, bl);
                mma3(s[j], qh[ks], ql[ks], bh, bl);
            }
        if (kt == ntiles - 1 && (N % kTile) != 0) {     // keys beyond N
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = kt * kTile + j * 8 + 2 * t;
                if (c >= N) { s[j][0] = -CUDART_INF_F; s[j][2] = -CUDART_INF_F; }
                if (c + 1 >= N) { s[j][1] = -CUDART_INF_F; s[j][3] = -CUDART_INF_F; }
            }
        }
        float mx_lo = -CUDART_INF_F, mx_hi = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx_lo = fmaxf(mx_lo, fmaxf(s[j][0], s[j][1]));
            mx_hi = fmaxf(mx_hi, fmaxf(s[j][2], s[j][3]));
        }
        mx_lo = quad_max(mx_lo); mx_hi = quad_max(mx_hi);
        const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
        const float c_lo = ex2f(m_lo - mn_lo), c_hi = ex2f(m_hi - mn_hi);
        m_lo = mn_lo; m_hi = mn_hi;
        float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = ex2f(s[j][0] - mn_lo); s[j][1] = ex2f(s[j][1] - mn_lo);
            s[j][2] = ex2f(s[j][2] - mn_hi); s[j][3] = ex2f(s[j][3] - mn_hi);
            sum_lo += s[j][0] + s[j][1]; sum_hi += s[j][2] + s[j][3];
        }
        l_lo = l_lo * c_lo + sum_lo; l_hi = l_hi * c_hi + sum_hi;
        float ot[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n) { ot[n][0] = ot[n][1] = ot[n][2] = ot[n][3] = 0.f; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t ph[4], pl[4];
            acc_to_a(s[j], ph, pl);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                uint32_t bh[2], bl[2];
                frag_rows<LD>(V, j * 8, n * 8, g, t, bh, bl);
                mma3(ot[n], ph, pl, bh, bl);
            }
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            oacc[n][0] = oacc[n][0] * c_lo + ot[n][0]; oacc[n][1] = oacc[n][1] * c_lo + ot[n][1];
            oacc[n][2] = oacc[n][2] * c_hi + ot[n][2]; oacc[n][3] = oacc[n][3] * c_hi + ot[n][3];
        }
    }
    cp_async_wait<0>();
    l_lo = quad_sum(l_lo); l_hi = quad_sum(l_hi);
    // ---- merge the four warps: [warp][16 rows] max / sum and [warp][16][HD] O through shared memory (tiles are dead)
    __syncthreads();
    float* sm_m = smem;                                         // [4][16]
    float* sm_l = smem + 64;                                    // [4][16]
    float* sm_o = smem + 128;                                   // [4][16][HD]
    if (t == 0) { sm_m[warp * 16 + g] = m_lo; sm_m[warp * 16 + g + 8] = m_hi; sm_l[warp * 16 + g] = l_lo; sm_l[warp * 16 + g + 8] = l_hi; }
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        *reinterpret_cast<float2*>(sm_o + (warp * 16 + g) * HD + n * 8 + 2 * t) = make_float2(oacc[n][0], oacc[n][1]);
        *reinterpret_cast<float2*>(sm_o + (warp * 16 + g + 8) * HD + n * 8 + 2 * t) = make_float2(oacc[n][2], oacc[n][3]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 16 * HD; i += 128) {
        const int r = i / HD, d = i - r * HD;
        if (q0 + r >= NQ) continue;
        const float M = fmaxf(fmaxf(sm_m[r], sm_m[16 + r]), fmaxf(sm_m[32 + r], sm_m[48 + r]));
        float L = 0.f, O = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float wgt = ex2f(sm_m[w * 16 + r] - M);      // exp2(-inf) = 0 for a warp that had no tile
            L += sm_l[w * 16 + r] * wgt;
            O += sm_o[(w * 16 + r) * HD + d] * wgt;
        }
        o[(int64_t)(q0 + r) * C + h * HD + d] = O / L;
        if (d == 0) lse[(int64_t)h * NQ + q0 + r] = (M + log2f(L)) * kLn2;
    }
}

// ---------------------------------------------------------------------------------------------------- backward: dQ
template <int HD>
__global__ void __launch_bounds__(kThreads)
attn_bwd_dq_tc_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, const float* __restrict__ lse,
                      const float* __restrict__ delta, int N, int C, float scale, float* __restrict__ dqkv) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int LD = HD + 4, KS = HD / 8, NT = HD / 8;
    extern __shared__ __align__(16) float smem[];
    constexpr int STAGE = 2 * kTile * LD;                       // K tile then V tile
    const int h = blockIdx.y, q0 = blockIdx.x * kRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;
    const bool ok_lo = r_lo < N, ok_hi = r_hi < N;
    const int64_t ld3 = 3ll * C;
    const float* kbase = qkv + C + h * HD;
    const float* vbase = qkv + 2 * C + h * HD;

    uint32_t qh[KS][4], ql[KS][4], gh[KS][4], gl[KS][4];
    load_a_rows<HD>(qkv + (int64_t)r_lo * ld3 + h * HD, qkv + (int64_t)r_hi * ld3 + h * HD, ok_lo, ok_hi, scale * kLog2e, t, qh, ql);
    load_a_rows<HD>(dO + (int64_t)r_lo * C + h * HD, dO + (int64_t)r_hi * C + h * HD, ok_lo, ok_hi, 1.0f, t, gh, gl);
    const float lse_lo = ok_lo ? lse[(int64_t)h * N + r_lo] * kLog2e : 0.f, lse_hi = ok_hi ? lse[(int64_t)h * N + r_hi] * kLog2e : 0.f;
    const float dl_lo = ok_lo ? delta[(int64_t)h * N + r_lo] : 0.f, dl_hi = ok_hi ? delta[(int64_t)h * N + r_hi] : 0.f;
    float dq[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) { dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f; }

    const int ntiles = (N + kTile - 1) / kTile;
    load_tile<HD>(smem, kbase, ld3, 0, N);
    load_tile<HD>(smem + kTile * LD, vbase, ld3, 0, N);
    cp_async_commit();
    for (int kt = 0; kt < ntiles; ++kt) {
        cp_async_wait<0>();
        __syncthreads();
        if (kt + 1 < ntiles) {
            float* nxt = smem + ((kt + 1) & 1) * STAGE;
            load_tile<HD>(nxt, kbase, ld3, (kt + 1) * kTile, N);
            load_tile<HD>(nxt + kTile * LD, vbase, ld3, (kt + 1) * kTile, N);
            cp_async_commit();
        }
        const float* K = smem + (kt & 1) * STAGE;
        const float* V = K + kTile * LD;
        float s[8][4], dp[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t bh[2], bl[2];
                frag_cols<LD>(K, j * 8, ks * 8, g, t, bh, bl);
                mma3(s[j], qh[ks], ql[ks], bh, bl);            // S = Q K^T (log2 units)
                frag_cols<LD>(V, j * 8, ks * 8, g, t, bh, bl);
                mma3(dp[j], gh[ks], gl[ks], bh, bl);           // dP = dO V^T
            }
        const bool tail = kt == ntiles - 1 && (N % kTile) != 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = kt * kTile + j * 8 + 2 * t;
            const bool v0 = !tail || c < N, v1 = !tail || c + 1 < N;
            // dS = P o (dP - delta)
            s[j][0] = v0 ? ex2f(s[j][0] - lse_lo) * (dp[j][0] - dl_lo) : 0.f;
            s[j][1] = v1 ? ex2f(s[j][1] - lse_lo) * (dp[j][1] - dl_lo) : 0.f;
            s[j][2] = v0 ? ex2f(s[j][2] - lse_hi) * (dp[j][2] - dl_hi) : 0.f;
            s[j][3] = v1 ? ex2f(s[j][3] - lse_hi) * (dp[j][3] - dl_hi) : 0.f;
        }
        float dt[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n) { dt[n][0] = dt[n][1] = dt[n][2] = dt[n][3] = 0.f; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t ah[4], al[4];
            acc_to_a(s[j], ah, al);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                uint32_t bh[2], bl[2];
                frag_rows<LD>(K, j * 8, n * 8, g, t, bh, bl);
                mma3(dt[n], ah, al, bh, bl);                   // dQ += dS K
            }
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) { dq[n][0] += dt[n][0]; dq[n][1] += dt[n][1]; dq[n][2] += dt[n][2]; dq[n][3] += dt[n][3]; }
    }
    if (ok_lo) {
#pragma unroll
        for (int n = 0; n < NT; ++n)
            *reinterpret_cast<float2*>(dqkv + (int64_t)r_lo * ld3 + h * HD + n * 8 + 2 * t) = make_float2(dq[n][0] * scale, dq[n][1] * scale);
    }
    if (ok_hi) {
#pragma unroll
        for (int n = 0; n < NT; ++n)
            *reinterpret_cast<float2*>(dqkv + (int64_t)r_hi * ld3 + h * HD + n * 8 + 2 * t) = make_float2(dq[n][2] * scale, dq[n][3] * scale);
    }
}

// ---------------------------------------------------------------------------------------------------- backward: dK, dV
template <int HD>
__global__ void __launch_bounds__(kThreads)
attn_bwd_dkv_tc_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, const float* __restrict__ lse,
                       const float* __restrict__ delta, int N, int C, float scale, float* __restrict__ dqkv) {
    pdl_wait();
    pdl_launch_dependents();
    constexpr int LD = HD + 4, KS = HD / 8, NT = HD / 8;
    constexpr int STAGE = 2 * kTile * LD + 2 * kTile;          // Q tile, dO tile, lse (log2 units), delta
    extern __shared__ __align__(16) float smem[];
    const int h = blockIdx.y, k0 = blockIdx.x * kRows;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int r_lo = k0 + warp * 16 + g, r_hi = r_lo + 8;       // key rows of this thread
    const bool ok_lo = r_lo < N, ok_hi = r_hi < N;
    const int64_t ld3 = 3ll * C;
    const float* qbase = qkv + h * HD;
    const float* gbase = dO + h * HD;

    // A operands: this warp's 16 keys (scaled into log2 units for the scores) and values
    uint32_t kh[KS][4], kl[KS][4], vh[KS][4], vl[KS][4];
    load_a_rows<HD>(qkv + (int64_t)r_lo * ld3 + C + h * HD, qkv + (int64_t)r_hi * ld3 + C + h * HD, ok_lo, ok_hi, scale * kLog2e, t, kh, kl);
    load_a_rows<HD>(qkv + (int64_t)r_lo * ld3 + 2 * C + h * HD, qkv + (int64_t)r_hi * ld3 + 2 * C + h * HD, ok_lo, ok_hi, 1.0f, t, vh, vl);
    float dk[NT][4], dv[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n) { dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f; }

    auto stage_load = [&](int st, int row0) {
        float* base = smem + st * STAGE;
        load_tile<HD>(base, qbase, ld3, row0, N);
        load_tile<HD>(base + kTile * LD, gbase, C, row0, N);
        if (threadIdx.x < kTile) {
            const int r = row0 + threadIdx.x;
            base[2 * kTile * LD + threadIdx.x] = r < N ? lse[(int64_t)h * N + r] * kLog2e : CUDART_INF_F;     // P = 0 beyond N
        } else {
            const int r = row0 + threadIdx.x - kTile;
            base[2 * kTile * LD + threadIdx.x] = r < N ? delta[(int64_t)h * N + r] : 0.f;
        }
    };
    const int ntiles = (N + kTile - 1) / kTile;
    stage_load(0, 0);
    cp_async_commit();
    for (int qt = 0; qt < ntiles; ++qt) {
        cp_async_wait<0>();
        __syncthreads();
        if (qt + 1 < ntiles) { stage_load((qt + 1) & 1, (qt + 1) * kTile); cp_async_commit(); }
        const float* Q = smem + (qt & 1) * STAGE;
        const float* G = Q + kTile * LD;
        const float* L = Q + 2 * kTile * LD;
        const float* Dl = L + kTile;
        float s[8][4], dp[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t bh[2], bl[2];
                frag_cols<LD>(Q, j * 8, ks * 8, g, t, bh, bl);
                mma3(s[j], kh[ks], kl[ks], bh, bl);            // S^T = K Q^T (log2 units)
                frag_cols<LD>(G, j * 8, ks * 8, g, t, bh, bl);
                mma3(dp[j], vh[ks], vl[ks], bh, bl);           // dP^T = V dO^T
            }
        // P^T = exp2(S^T - lse[query]);  dS^T = P^T o (dP^T - delta[query]);  the thread's columns are queries j*8 + 2t, +1
        float dvt[NT][4], dkt[NT][4];
#pragma unroll
        for (int n = 0; n < NT; ++n) { dvt[n][0] = dvt[n][1] = dvt[n][2] = dvt[n][3] = 0.f; dkt[n][0] = dkt[n][1] = dkt[n][2] = dkt[n][3] = 0.f; }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 l2 = *reinterpret_cast<const float2*>(L + j * 8 + 2 * t);
            const float2 d2 = *reinterpret_cast<const float2*>(Dl + j * 8 + 2 * t);
            float p[4] = {ex2f(s[j][0] - l2.x), ex2f(s[j][1] - l2.y), ex2f(s[j][2] - l2.x), ex2f(s[j][3] - l2.y)};
            float ds[4] = {p[0] * (dp[j][0] - d2.x), p[1] * (dp[j][1] - d2.y), p[2] * (dp[j][2] - d2.x), p[3] * (dp[j][3] - d2.y)};
            uint32_t ph[4], pl[4], sh[4], sl[4];
            acc_to_a(p, ph, pl);
            acc_to_a(ds, sh, sl);
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                uint32_t bh[2], bl[2];
                frag_rows<LD>(G, j * 8, n * 8, g, t, bh, bl);
                mma3(dvt[n], ph, pl, bh, bl);                  // dV += P^T dO
                frag_rows<LD>(Q, j * 8, n * 8, g, t, bh, bl);
                mma3(dkt[n], sh, sl, bh, bl);                  // dK += dS^T Q
            }
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            dk[n][0] += dkt[n][0]; dk[n][1] += dkt[n][1]; dk[n][2] += dkt[n][2]; dk[n][3] += dkt[n][3];
            dv[n][0] += dvt[n][0]; dv[n][1] += dvt[n][1]; dv[n][2] += dvt[n][2]; dv[n][3] += dvt[n][3];
        }
    }
    if (ok_lo) {
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            *reinterpret_cast<float2*>(dqkv + (int64_t)r_lo * ld3 + C + h * HD + n * 8 + 2 * t) = make_float2(dk[n][0] * scale, dk[n][1] * scale);
            *reinterpret_cast<float2*>(dqkv + (int64_t)r_lo * ld3 + 2 * C + h * HD + n * 8 + 2 * t) = make_float2(dv[n][0], dv[n][1]);
        }
    }
    if (ok_hi) {
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            *reinterpret_cast<float2*>(dqkv + (int64_t)r_hi * ld3 + C + h * HD + n * 8 + 2 * t) = make_float2(dk[n][2] * scale, dk[n][3] * scale);
            *reinterpret_cast<float2*>(dqkv + (int64_t)r_hi * ld3 + 2 * C + h * HD + n * 8 + 2 * t) = make_float2(dv[n][2], dv[n][3]);
        }
    }
}

template <int HD> constexpr size_t kv_smem() { return (size_t)4 * kTile * (HD + 4) * sizeof(float); }
template <int HD> constexpr size_t dkv_smem() { return (size_t)2 * (2 * kTile * (HD + 4) + 2 * kTile) * sizeof(float); }

template <int HD, int WARPS>
static void fwd_w(const float* qkv, int N, int C, int H, float scale, float* o, float* lse, cudaStream_t st, int q_start,
                  int q_stride, int NQ) {
    ensure_dyn_smem(reinterpret_cast<const void*>(attn_fwd_tc_kernel<HD, WARPS>), (int)kv_smem<HD>());
    launch_pdl(attn_fwd_tc_kernel<HD, WARPS>, dim3((NQ + 16 * WARPS - 1) / (16 * WARPS), H), dim3(32 * WARPS), kv_smem<HD>(), st, qkv,
               N, C, scale, o, lse, q_start, q_stride, NQ);
}
// Keys split over the warps from four key tiles up (MOMA_B200_ATTN_SPLITKV=0: never).  Measured on the C3 step: the three
// N = 512 forward launches 41.8 -> 34.0 us, step 0.1973 -> 0.1938 ms; owned-rows attention at 8 GPUs 131 -> 74 us.
static bool use_splitkv(int N, int NQ) {
    static const bool off = [] { const char* e = getenv("MOMA_B200_ATTN_SPLITKV"); return e != nullptr && e[0] == '0'; }();
    (void)NQ;
    return !off && N >= 4 * kTile;
}
template <int HD>
static void fwd(const float* qkv, int N, int C, int H, float scale, float* o, float* lse, cudaStream_t st, int q_start,
                int q_stride, int NQ) {
    if constexpr (HD <= 32) {                                   // (head_dim 64: four warps' private tiles exceed shared memory)
        if (use_splitkv(N, NQ)) {
            constexpr size_t smem = (size_t)4 * 2 * 2 * kTile * (HD + 4) * sizeof(float);      // 4 warps x 2 stages x (K, V)
            ensure_dyn_smem(reinterpret_cast<const void*>(attn_fwd_splitkv_kernel<HD>), (int)smem);
            launch_pdl(attn_fwd_splitkv_kernel<HD>, dim3((NQ + 15) / 16, H), dim3(128), smem, st, qkv, N, C, scale, o, lse, q_start,
                       q_stride, NQ);
            return;
        }
    }
    // 4 warps (64 query rows) per CTA.  Smaller CTAs (2 / 1 warps, more CTAs when rows x heads < SMs) were measured at
    // 8 GPUs on the owned-rows attention (512 rows x 4096 keys): 128 two-warp CTAs took 219 us where 64 four-warp CTAs take
    // 77 us next to the rest of the step -- every CTA streams all K/V tiles, and fewer, fatter CTAs disturb fewer SMs.
    // MOMA_B200_ATTN_WARPS=2|1 selects them (A/B switch, read once).
    static const int warps = [] { const char* e = getenv("MOMA_B200_ATTN_WARPS"); int v = e ? atoi(e) : 4; return (v == 1 || v == 2) ? v : 4; }();
    if (warps == 4) fwd_w<HD, 4>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ);
    else if (warps == 2) fwd_w<HD, 2>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ);
    else fwd_w<HD, 1>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ);
}
template <int HD>
static void bwd(const float* qkv, const float* dO, const float* lse, const float* delta, int N, int C, int H, float scale,
                float* dqkv, cudaStream_t st, cudaStream_t st2) {
    ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dq_tc_kernel<HD>), (int)kv_smem<HD>());
    ensure_dyn_smem(reinterpret_cast<const void*>(attn_bwd_dkv_tc_kernel<HD>), (int)dkv_smem<HD>());
    const dim3 grid((N + kRows - 1) / kRows, H);
    launch_pdl(attn_bwd_dq_tc_kernel<HD>, grid, dim3(kThreads), kv_smem<HD>(), st, qkv, dO, lse, delta, N, C, scale, dqkv);
    launch_pdl(attn_bwd_dkv_tc_kernel<HD>, grid, dim3(kThreads), dkv_smem<HD>(), st2, qkv, dO, lse, delta, N, C, scale, dqkv);
}

}  // namespace atc

// MOMA_B200_ATTN=simt: the FP32 CUDA-core kernels of attn.cu for every head size (A/B switch, read once)
bool attn_tc_supported(int hd) {
    static const bool simt = [] { const char* e = getenv("MOMA_B200_ATTN"); return e != nullptr && std::string(e) == "simt"; }();
    static const bool no64 = [] { const char* e = getenv("MOMA_B200_ATTN_TC64"); return e != nullptr && e[0] == '0'; }();
    return !simt && (hd == 8 || hd == 16 || hd == 32 || (hd == 64 && !no64));
}
void attn_tc_fwd(const float* qkv, int N, int C, int H, float scale, float* o, float* lse, cudaStream_t st, int q_start,
                 int q_stride, int NQ) {
    note_flops(1, 4.0 * NQ * N * C);
    switch (C / H) {
        case 8: atc::fwd<8>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ); break;
        case 16: atc::fwd<16>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ); break;
        case 64: atc::fwd<64>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ); break;
        default: atc::fwd<32>(qkv, N, C, H, scale, o, lse, st, q_start, q_stride, NQ); break;
    }
}
void attn_tc_bwd(const float* qkv, const float* dO, const float* lse, const float* delta, int N, int C, int H, float scale,
                 float* dqkv, cudaStream_t st, cudaStream_t st2) {
    note_flops(1, 8.0 * N * N * C);
    switch (C / H) {
        case 8: atc::bwd<8>(qkv, dO, lse, delta, N, C, H, scale, dqkv, st, st2); break;
        case 16: atc::bwd<16>(qkv, dO, lse, delta, N, C, H, scale, dqkv, st, st2); break;
        case 64: atc::bwd<64>(qkv, dO, lse, delta, N, C, H, scale, dqkv, st, st2); break;
        default: atc::bwd<32>(qkv, dO, lse, delta, N, C, H, scale, dqkv, st, st2); break;
    }
}

}  // namespace moma
