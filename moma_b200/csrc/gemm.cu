// Small-matrix GEMM for the projection heads and the attention projections of the MoMA criterion:
//   C[M, N] = act( A · B^T + bias ),  A[m, k] = A[m * a_rs + k * a_cs],  B[n, k] = B[n * b_rs + k * b_cs].
//
// The matrices of this path are tiny (M = batch = 256 rows, N, K <= 2048), so a launch is bound by latency and
// by how many SMs it reaches, not by arithmetic: 32 x 32 output tiles (every tile its own CTA), split along K
// when the tile grid would not fill the 148 SMs, the next K-slab in flight in registers while the current one
// is multiplied.  The products run on the tensor cores as 3xTF32 (a = a_hi + a_lo, both TF32; a·b ~=
// a_lo·b_hi + a_hi·b_lo + a_hi·b_hi accumulated in fp32), which keeps fp32-level accuracy (the dropped a_lo·b_lo term
// is 2^-22 relative) -- a single TF32 pass is NOT accurate enough for the student's gradients (scripts/tf32_heads_error.py).
// Warp-level mma.sync is deliberate here: a 32 x 32 x K tile has no use for a 128-row tcgen05 accumulator.
//
// Split-K is deterministic: every split writes its partial tile to the workspace, takes a ticket, and the last
// arrival sums the partials in split order, applies bias / ReLU and writes C (the ticket counter resets itself).
#include <cstdlib>
#include "common.cuh"

namespace moma {

namespace {

constexpr int kTM = 32, kTN = 32, kTK = 32, kThreads = 128;
constexpr int kLdK = kTK + 4;        // [row][k] tiles: 36-float rows  -> fragment reads hit banks 4g + t
constexpr int kLdR = kTM + 8;        // [k][row] tiles: 40-float rows  -> fragment reads hit banks 8t + g
constexpr int kTile = kTK * kLdR;    // floats per operand tile (the larger of the two layouts)
constexpr int kStages = 3;
constexpr int kMaxSplits = 16;

// x = hi + lo with hi = x rounded to TF32 (10-bit mantissa, round-to-nearest on the bit pattern: integer add of half an
// ulp, then mask) and lo = x - hi (exact in fp32).  The tensor core reads only the top 19 bits of lo, i.e. truncates it:
// |error| <= 2^-21 |x|, unbiased because lo's sign is.  Three ALU/FMA-pipe instructions per element.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// operand access pattern of a [rows, K] operand
enum : int { kVecK = 0,    // k contiguous (cs == 1), 16-byte aligned rows: 16-byte copies along k, tile kept [row][k]
             kVecR = 1,    // row index contiguous (rs == 1): 16-byte copies along the rows, tile kept [k][row]
             kScalar = 2 };// anything else: 4-byte copies, tile kept [row][k]

struct Operand {
    const float* p;
    int64_t rs, cs;
    int rows, mode;
};

// global -> shared, asynchronously (zero-filled outside the matrix): one 32 x 32 tile, 8 floats per thread
__device__ __forceinline__ void issue_tile(float* dst, const float* __restrict__ src, const Operand& op, int row0, int k0, int K) {
    const int tid = threadIdx.x;
    if (op.mode == kVecK) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * kThreads, rr = idx >> 3, kv = (idx & 7) * 4;
            const bool ok = row0 + rr < op.rows && k0 + kv < K;
            cp_async16(dst + rr * kLdK + kv, ok ? src + (int64_t)(row0 + rr) * op.rs + k0 + kv : src, ok);
        }
    } else if (op.mode == kVecR) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * kThreads, kk = idx >> 3, rv = (idx & 7) * 4;
            const bool ok = k0 + kk < K && row0 + rv < op.rows;
            cp_async16(dst + kk * kLdR + rv, ok ? src + (int64_t)(k0 + kk) * op.cs + row0 + rv : src, ok);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * kThreads, rr = idx >> 5, kk = idx & 31;
            const bool ok = row0 + rr < op.rows && k0 + kk < K;
            cp_async4(dst + rr * kLdK + kk, ok ? src + (int64_t)(row0 + rr) * op.rs + (int64_t)(k0 + kk) * op.cs : src, ok);
        }
    }
}

struct GemmParams {
    Operand a, b;
    const float* a_mask;       // optional, indexed like A: elements with mask <= 0 read as 0 (ReLU backward)
    const float* bias;
    float* C;
    __nv_bfloat16* C_bf16;     // optional second output: the same values rounded to bf16, dense [M, N]
    int64_t ldc;
    int M, N, K, relu, splits;
    float* partial;            // [splits, M, N] when splits > 1
    unsigned int* tickets;     // one per output tile, zero between launches
};

// Per-thread state of one operand's tile copies in the two vector modes: everything that does not change from slab to
// slab (source offset, shared-memory offset, the row bound) is computed once, so a slab costs two pointer adds, two
// compares and two cp.async per operand instead of re-deriving 64-bit addresses through the mode dispatch.
struct TileLoader {
    long long off[2];      // element offset of this thread's two 16-byte vectors at slab 0
    long long kstep;       // element offset between consecutive slabs
    int dst[2];            // float offset inside the shared-memory tile
    int kidx[2];           // k index (inside the slab) of the vector's first element
    bool ok[2];            // row bound (static)
};
__device__ __forceinline__ void loader_init(TileLoader& L, const Operand& op, int row0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int idx = threadIdx.x + i * kThreads, o = idx >> 3, v = (idx & 7) * 4;
        if (op.mode == kVecK) {          // outer = row, vector along k
            L.ok[i] = row0 + o < op.rows; L.kidx[i] = v; L.dst[i] = o * kLdK + v;
            L.off[i] = (long long)(row0 + o) * op.rs + v;
        } else {                          // kVecR: outer = k, vector along the rows
            L.ok[i] = row0 + v < op.rows; L.kidx[i] = o; L.dst[i] = o * kLdR + v;
            L.off[i] = (long long)o * op.cs + row0 + v;
        }
    }
    L.kstep = op.mode == kVecK ? (long long)kTK : (long long)kTK * op.cs;
}
__device__ __forceinline__ void loader_issue(const TileLoader& L, float* tile, const float* __restrict__ base, int kt, int K) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const bool ok = L.ok[i] && kt * kTK + L.kidx[i] < K;
        cp_async16(tile + L.dst[i], ok ? base + L.off[i] + kt * L.kstep : base, ok);
    }
}

// SCALAR: at least one operand needs element-wise copies (odd shapes / unaligned views): generic, slower addressing
template <bool MASK, bool SCALAR>
__global__ void __launch_bounds__(kThreads)
gemm3xtf32_kernel(const GemmParams p) {
    __shared__ __align__(16) float sA[kStages][kTile];
    __shared__ __align__(16) float sB[kStages][kTile];
    __shared__ __align__(16) float sM[MASK ? kStages : 1][MASK ? kTile : 4];
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 16;
    const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
    const int ktiles = (p.K + kTK - 1) / kTK;
    const int kt0 = (int)((long long)ktiles * blockIdx.z / p.splits);
    const int kt1 = (int)((long long)ktiles * (blockIdx.z + 1) / p.splits);
    // shared-memory strides of the two tile layouts
    const int a_sr = p.a.mode == kVecR ? 1 : kLdK, a_sk = p.a.mode == kVecR ? kLdR : 1;
    const int b_sr = p.b.mode == kVecR ? 1 : kLdK, b_sk = p.b.mode == kVecR ? kLdR : 1;
    const int a_base = (wm + g) * a_sr + t * a_sk;
    const int b_base = (wn + g) * b_sr + t * b_sk;

    pdl_wait();                     // operands may come from the kernel launched just before this one
    TileLoader la, lb;
    if (!SCALAR) { loader_init(la, p.a, m0); loader_init(lb, p.b, n0); }
    auto issue = [&](int stage, int kt) {
        if (SCALAR) {
            issue_tile(sA[stage], p.a.p, p.a, m0, kt * kTK, p.K);
            issue_tile(sB[stage], p.b.p, p.b, n0, kt * kTK, p.K);
            if (MASK) issue_tile(sM[stage], p.a_mask, p.a, m0, kt * kTK, p.K);
        } else {
            loader_issue(la, sA[stage], p.a.p, kt, p.K);
            loader_issue(lb, sB[stage], p.b.p, kt, p.K);
            if (MASK) loader_issue(la, sM[stage], p.a_mask, kt, p.K);
        }
    };
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (kt0 + s < kt1) issue(s, kt0 + s);
        cp_async_commit();
    }

    float acc[2][4] = {};
    for (int kt = kt0; kt < kt1; ++kt) {
        const int st = (kt - kt0) % kStages;
        cp_async_wait<kStages - 2>();
        __syncthreads();                      // slab kt has landed for everyone; slab kt - 1's buffer is free again
        if (kt + kStages - 1 < kt1) issue((kt - kt0 + kStages - 1) % kStages, kt + kStages - 1);
        cp_async_commit();
        const float* A = sA[st];
        const float* Bt = sB[st];
        const float* Mk = sM[MASK ? st : 0];
        // The tensor core adds into its accumulator with truncation, so a long chain drifts (measured 2.5e-6
        // relative at K = 512 against 2e-7 for slab-long chains): accumulate one slab, then add it to the running
        // sum with an ordinary round-to-nearest FADD.
        // Separate accumulators for the three product terms: six independent MMA chains per warp instead of two
        // three-deep dependent ones (mma.sync latency ~50 clk against ~9 clk issue).
        float plh[2][4] = {}, phl[2][4] = {}, phh[2][4] = {};
#pragma unroll
        for (int kk = 0; kk < kTK; kk += 8) {
            uint32_t ah[4], al[4];
            const int ia = a_base + kk * a_sk;
            float x[4] = {A[ia], A[ia + 8 * a_sr], A[ia + 4 * a_sk], A[ia + 8 * a_sr + 4 * a_sk]};
            if (MASK) {
                x[0] = Mk[ia] > 0.f ? x[0] : 0.f;                       x[1] = Mk[ia + 8 * a_sr] > 0.f ? x[1] : 0.f;
                x[2] = Mk[ia + 4 * a_sk] > 0.f ? x[2] : 0.f;            x[3] = Mk[ia + 8 * a_sr + 4 * a_sk] > 0.f ? x[3] : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) split_tf32(x[e], ah[e], al[e]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int ib = b_base + 8 * j * b_sr + kk * b_sk;
                uint32_t bh[2], bl[2];
                split_tf32(Bt[ib], bh[0], bl[0]);
                split_tf32(Bt[ib + 4 * b_sk], bh[1], bl[1]);
                mma_tf32(plh[j], al, bh);
                mma_tf32(phl[j], ah, bl);
                mma_tf32(phh[j], ah, bh);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] += (plh[j][e] + phl[j][e]) + phh[j][e];     // small terms first
    }
    cp_async_wait<0>();
    pdl_launch_dependents();        // the successor may start launching while the epilogue (and split-K fix-up) runs

    // accumulator fragment: acc[j][0..1] -> row wm + g, cols wn + 8j + 2t (+1); acc[j][2..3] -> row + 8
    auto finish = [&](int r, int c, float v0, float v1) {
        if (r >= p.M) return;
        if (p.bias) { if (c < p.N) v0 += p.bias[c]; if (c + 1 < p.N) v1 += p.bias[c + 1]; }
        if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
        float* dst = p.C + (int64_t)r * p.ldc + c;
        if (c + 1 < p.N && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
        else { if (c < p.N) dst[0] = v0; if (c + 1 < p.N) dst[1] = v1; }
        if (p.C_bf16) {
            __nv_bfloat16* d16 = p.C_bf16 + (int64_t)r * p.N + c;
            if (c + 1 < p.N && ((reinterpret_cast<uintptr_t>(d16) & 3) == 0)) *reinterpret_cast<__nv_bfloat162*>(d16) = __floats2bfloat162_rn(v0, v1);
            else { if (c < p.N) d16[0] = __float2bfloat16_rn(v0); if (c + 1 < p.N) d16[1] = __float2bfloat16_rn(v1); }
        }
    };
    if (p.splits == 1) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = n0 + wn + 8 * j + 2 * t;
            finish(m0 + wm + g, c, acc[j][0], acc[j][1]);
            finish(m0 + wm + g + 8, c, acc[j][2], acc[j][3]);
        }
        return;
    }
    // ---- split-K: publish the partial tile, last arrival reduces in split order
    {
        float* mine = p.partial + (int64_t)blockIdx.z * p.M * p.N;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = n0 + wn + 8 * j + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = m0 + wm + g + 8 * h;
                if (r < p.M) {
                    if (c < p.N) __stcg(mine + (int64_t)r * p.N + c, acc[j][2 * h]);
                    if (c + 1 < p.N) __stcg(mine + (int64_t)r * p.N + c + 1, acc[j][2 * h + 1]);
                }
            }
        }
    }
    __threadfence();
    __syncthreads();
    const unsigned tile = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) s_last = (atomicAdd(p.tickets + tile, 1u) == (unsigned)p.splits - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int c = n0 + wn + 8 * j + 2 * t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = m0 + wm + g + 8 * h;
            if (r >= p.M) continue;
            float v0 = 0.f, v1 = 0.f;
            for (int s = 0; s < p.splits; ++s) {
                const float* src = p.partial + ((int64_t)s * p.M + r) * p.N + c;
                if (c < p.N) v0 += __ldcg(src);
                if (c + 1 < p.N) v1 += __ldcg(src + 1);
            }
            finish(r, c, v0, v1);
        }
    }
    if (tid == 0) p.tickets[tile] = 0u;
}

// column sums of a (masked) [rows, cols] matrix: bias gradients.  8 columns per CTA (cols / 8 CTAs), 32 row phases:
// a thread sums every 32nd row of one column with its loads unrolled (one memory round trip for the usual 256-row batch),
// then the 32 phases are added in a fixed order (deterministic).
constexpr int kCsCols = 8, kCsPhases = 32;
__global__ void __launch_bounds__(256)
colsum_masked_kernel(const float* __restrict__ X, const float* __restrict__ mask, int rows, int cols, float* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float red[kCsPhases][kCsCols + 1];
    const int cl = threadIdx.x & (kCsCols - 1), ph = threadIdx.x >> 3;
    const int c = blockIdx.x * kCsCols + cl;
    float s = 0.f;
    if (c < cols) {
        int r = ph;
        for (; r + 7 * kCsPhases < rows; r += 8 * kCsPhases) {
            float v[8], m[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = X[(int64_t)(r + u * kCsPhases) * cols + c];
            if (mask) {
#pragma unroll
                for (int u = 0; u < 8; ++u) m[u] = mask[(int64_t)(r + u * kCsPhases) * cols + c];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = m[u] > 0.f ? v[u] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; r < rows; r += kCsPhases) {
            const float v = X[(int64_t)r * cols + c];
            s += (mask == nullptr || mask[(int64_t)r * cols + c] > 0.f) ? v : 0.f;
        }
    }
    red[ph][cl] = s;
    __syncthreads();
    if (threadIdx.x < kCsCols && c < cols) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < kCsPhases; ++i) tot += red[i][threadIdx.x];
        out[c] = tot;
    }
}

int operand_mode(const float* p, const float* mask, int64_t rs, int64_t cs, int rows, int K) {
    const bool al = aligned16(p) && (mask == nullptr || aligned16(mask));
    if (cs == 1 && al && rs % 4 == 0 && K % 4 == 0) return kVecK;
    if (rs == 1 && al && cs % 4 == 0 && rows % 4 == 0) return kVecR;
    return kScalar;
}

}  // namespace

int gemm_splits(int M, int N, int K) {
    const int tiles = ((M + kTM - 1) / kTM) * ((N + kTN - 1) / kTN);
    const int ktiles = (K + kTK - 1) / kTK;
    // CTAs aimed at per SM (MOMA_B200_GEMM_OVERSUB, read once; default 2: measured -3 % on the C2 step against 1, 3 gives nothing more)
    static const int oversub = [] { const char* e = getenv("MOMA_B200_GEMM_OVERSUB"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
    // deep contractions (K >= 1024: the teacher head's second layer, 2048 -> D) have few output tiles and many slabs:
    // twice the CTAs again, the extra fix-up reads are L2 hits
    const int target = sm_count() * oversub * (ktiles >= 32 ? 2 : 1);
    if (tiles * 5 >= target * 4) return 1;                     // >= 0.8 wave already
    int s = target / tiles;
    s = s < ktiles / 4 ? s : ktiles / 4;                       // >= 4 K-slabs per split: the fix-up costs two L2 round trips
    s = s < kMaxSplits ? s : kMaxSplits;
    return s < 1 ? 1 : s;
}

static size_t gemm_simt_workspace_bytes(int M, int N, int K) {
    const int s = gemm_splits(M, N, K);
    if (s == 1) return 0;
    const size_t tiles = (size_t)((M + kTM - 1) / kTM) * ((N + kTN - 1) / kTN);
    return ((tiles * sizeof(unsigned) + 255) / 256) * 256 + (size_t)s * M * N * sizeof(float);
}
// either kernel may end up running the shape (the choice also depends on pointers and strides): size for both
size_t gemm_workspace_bytes(int M, int N, int K) {
    const size_t a = gemm_simt_workspace_bytes(M, N, K), b = M >= 128 ? gemm_tc_workspace_bytes(M, N, K) : 0;
    return a > b ? a : b;
}

// workspace == nullptr forces a single split (no workspace needed).  The first `tiles` words of the workspace are the
// ticket counters: they must be zero before the first use and are left zero by every launch.
int gemm_nt(const float* A, const float* a_mask, int64_t a_rs, int64_t a_cs, const float* Bm, int64_t b_rs, int64_t b_cs,
            const float* bias, float* C, int64_t ldc, int M, int N, int K, int relu, void* workspace, size_t workspace_bytes,
            cudaStream_t st, void* c_bf16) {
    // tcgen05 kernel (gemm_tc.cu) for tall-enough outputs with TMA-compatible operands
    if (c_bf16 == nullptr && (a_cs == 1 || a_rs == 1) && (b_cs == 1 || b_rs == 1)) {
        const int a_mn = (a_cs != 1), b_mn = (b_cs != 1);
        const long long lda = a_mn ? a_cs : a_rs, ldb = b_mn ? b_cs : b_rs;
        if (!(a_mn && !b_mn) && gemm_tc_supported(A, lda, Bm, ldb, M, N, K) && (a_mask == nullptr || aligned16(a_mask)) &&
            (a_mn ? M % 4 == 0 : K % 4 == 0) && (b_mn ? N % 4 == 0 : K % 4 == 0))     // float4 along each contiguous dimension
            return gemm_tc(A, lda, a_mn, a_mask, lda, Bm, ldb, b_mn, bias, C, ldc, M, N, K, relu, workspace, workspace_bytes, st);
    }
    note_flops(0, 2.0 * M * N * K);
    GemmParams p;
    p.a = Operand{A, a_rs, a_cs, M, operand_mode(A, a_mask, a_rs, a_cs, M, K)};
    p.b = Operand{Bm, b_rs, b_cs, N, operand_mode(Bm, nullptr, b_rs, b_cs, N, K)};
    p.a_mask = a_mask; p.bias = bias; p.C = C; p.C_bf16 = static_cast<__nv_bfloat16*>(c_bf16); p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.relu = relu;
    p.splits = 1; p.partial = nullptr; p.tickets = nullptr;
    if (workspace != nullptr) {
        const int s = gemm_splits(M, N, K);
        if (s > 1) {
            MOMA_REQUIRE(workspace_bytes >= gemm_simt_workspace_bytes(M, N, K) && aligned16(workspace), MOMA_ERR_WORKSPACE,
                         "gemm: workspace too small or unaligned");
            const size_t tiles = (size_t)((M + kTM - 1) / kTM) * ((N + kTN - 1) / kTN);
            p.splits = s;
            p.tickets = static_cast<unsigned int*>(workspace);
            p.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + ((tiles * sizeof(unsigned) + 255) / 256) * 256);
        }
    }
    const dim3 grid((N + kTN - 1) / kTN, (M + kTM - 1) / kTM, p.splits);
    const bool scalar = p.a.mode == kScalar || p.b.mode == kScalar;
    if (scalar) {
        if (a_mask) launch_pdl(gemm3xtf32_kernel<true, true>, grid, dim3(kThreads), 0, st, p);
        else launch_pdl(gemm3xtf32_kernel<false, true>, grid, dim3(kThreads), 0, st, p);
    } else {
        if (a_mask) launch_pdl(gemm3xtf32_kernel<true, false>, grid, dim3(kThreads), 0, st, p);
        else launch_pdl(gemm3xtf32_kernel<false, false>, grid, dim3(kThreads), 0, st, p);
    }
    MOMA_CUDA_LAUNCH_CHECK("gemm3xtf32");
    return MOMA_OK;
}

void colsum_masked(const float* X, const float* mask, int rows, int cols, float* out, cudaStream_t st) {
    launch_pdl(colsum_masked_kernel, dim3((cols + kCsCols - 1) / kCsCols), dim3(256), 0, st, X, mask, rows, cols, out);
}

}  // namespace moma

using namespace moma;

// ----------------------------------------------------------------------------- C ABI: Linear (+ReLU)
// Replaces the nn.Linear / nn.ReLU pairs of the projection heads (MoMA/criterion_moco_att.py:254-305).
extern "C" __attribute__((visibility("default"))) size_t moma_linear_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    if (M <= 0 || N <= 0 || K <= 0 || M > (1 << 24) || N > (1 << 24) || K > (1 << 24)) return 0;
    // forward: [M,N] over K; backward: dX [M,K] over N and dW [N,K] over M run concurrently on disjoint halves
    const size_t f = gemm_workspace_bytes((int)M, (int)N, (int)K);
    const size_t b = gemm_workspace_bytes((int)M, (int)K, (int)N) + gemm_workspace_bytes((int)N, (int)K, (int)M);
    return (f > b ? f : b) + 256;
}

extern "C" __attribute__((visibility("default"))) int moma_linear_fwd(
    const float* x, const float* w, const float* b, int64_t M, int64_t N, int64_t K, int relu, float* y,
    void* workspace, size_t workspace_bytes, moma_stream_t stream) {
    MOMA_REQUIRE(x && w && y, MOMA_ERR_INVALID, "linear_fwd: null pointer");
    MOMA_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1 << 24) && N < (1 << 24) && K < (1 << 24), MOMA_ERR_INVALID,
                 "linear_fwd: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    int rc = gemm_nt(x, nullptr, K, 1, w, K, 1, b, y, N, (int)M, (int)N, (int)K, relu, workspace, workspace_bytes, as_stream(stream));
    if (rc != MOMA_OK) return rc;
    note_launches(1);
    return MOMA_OK;
}

namespace {
struct LinStreams { cudaStream_t s1 = nullptr; cudaEvent_t fork = nullptr, join = nullptr; bool ok = false; };
LinStreams& lin_streams() {
    static thread_local LinStreams b;
    static thread_local bool init = false;
    if (!init) {
        init = true;
        bool good = cudaStreamCreateWithFlags(&b.s1, cudaStreamNonBlocking) == cudaSuccess;
        good = good && cudaEventCreateWithFlags(&b.fork, cudaEventDisableTiming) == cudaSuccess;
        good = good && cudaEventCreateWithFlags(&b.join, cudaEventDisableTiming) == cudaSuccess;
        b.ok = good;
        if (!good) cudaGetLastError();
    }
    return b;
}
}  // namespace

// y is the forward OUTPUT (post-ReLU when relu != 0; used as the ReLU mask, may be NULL when relu == 0).
// Any of grad_x / grad_w / grad_b may be NULL.  dX runs on `stream`, dW / db on an internal side stream that is
// forked from and joined back into `stream` (parallel branches under graph capture).
extern "C" __attribute__((visibility("default"))) int moma_linear_bwd(
    const float* x, const float* w, const float* y, const float* grad_y, int64_t M, int64_t N, int64_t K, int relu,
    float* grad_x, float* grad_w, float* grad_b, void* workspace, size_t workspace_bytes, moma_stream_t stream) {
    MOMA_REQUIRE(x && w && grad_y && (!relu || y), MOMA_ERR_INVALID, "linear_bwd: null pointer");
    MOMA_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1 << 24) && N < (1 << 24) && K < (1 << 24), MOMA_ERR_INVALID,
                 "linear_bwd: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    cudaStream_t st = as_stream(stream);
    const float* mask = relu ? y : nullptr;
    const int m = (int)M, n = (int)N, k = (int)K;
    char* ws = static_cast<char*>(workspace);
    const size_t need_x = gemm_workspace_bytes(m, k, n), need_w = gemm_workspace_bytes(n, k, m);
    const bool have_ws = ws != nullptr && workspace_bytes >= need_x + need_w + 256 && aligned16(ws);
    void* ws_x = have_ws && need_x ? ws : nullptr;
    void* ws_w = have_ws && need_w ? ws + ((need_x + 255) / 256) * 256 : nullptr;
    LinStreams& ls = lin_streams();
    const bool side = ls.ok && (grad_w || grad_b) && grad_x;
    cudaStream_t s1 = side ? ls.s1 : st;
    if (side) { cudaEventRecord(ls.fork, st); cudaStreamWaitEvent(s1, ls.fork, 0); }
    int launches = 0, rc = MOMA_OK;
    // side branch: the short bias-gradient kernel first, so the branch ends with the long GEMM
    if (grad_b) { colsum_masked(grad_y, mask, m, n, grad_b, s1); ++launches; }
    // dW[n, k] = sum_m g[m, n] x[m, k]
    if (grad_w) {
        rc = gemm_nt(grad_y, mask, 1, N, x, 1, K, nullptr, grad_w, K, n, k, m, 0, ws_w, need_w, s1);
        if (rc != MOMA_OK) return rc;
        ++launches;
    }
    // dX[m, k] = sum_n g[m, n] w[n, k]
    if (grad_x) {
        rc = gemm_nt(grad_y, mask, N, 1, w, 1, K, nullptr, grad_x, K, m, k, n, 0, ws_x, need_x, st);
        if (rc != MOMA_OK) return rc;
        ++launches;
    }
    if (side) { cudaEventRecord(ls.join, s1); cudaStreamWaitEvent(st, ls.join, 0); }
    MOMA_CUDA_LAUNCH_CHECK("linear_bwd");
    note_launches(launches);
    return MOMA_OK;
}
