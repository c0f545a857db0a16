// 3xTF32 GEMM on tcgen05 tensor cores for the Linear layers of the projection heads and the attention projections:
//     C[M, N] = act( A[M, K] . B[N, K]^T + bias )            (forward: A, B row-major with K contiguous)
//     C[M, N] = ( A[M, K] o mask ) . B[K, N]                   (d loss / d input: B with N contiguous)
//     C[M, N] = ( A[K, M] o mask )^T . B[K, N]                 (d loss / d weight: both operands with the contraction outermost)
// Replaces the same lines as csrc/gemm.cu (MoMA/criterion_moco_att.py:254-305 and their autograd) for outputs of at least
// 128 rows.  The warp-level kernel of gemm.cu spends ~90 % of its issue slots on fragment loads and hi / lo splitting
// (ncu: tensor pipe 13 % active); here
//   * eight producer warps read fp32 tiles [128 x 32] of A and [64 x 32] of B from global memory (the next two K-slabs are
//     in flight while the current one is processed), apply the ReLU mask, split every value into its TF32-rounded "hi" and the
//     exact remainder "lo" (x = hi + lo) IN REGISTERS and store the two planes into shared memory in the tensor core's
//     canonical layout (128-byte swizzle for K-major operands, 128-byte swizzle with 32-byte base for MN-major ones -- the
//     only layout tcgen05 reads MN-major TF32 from).  Shared memory is written once and read only by the tensor core;
//   * one thread issues, per 8-wide k-step, two tcgen05.mma kind::tf32: lo.hi (M128 N64 K8) and hi.(hi | lo) (M128 N128 K8:
//     the hi and lo planes of B are adjacent, so A_hi is read once for both products), each product into its OWN TMEM
//     accumulator (the tensor core adds with truncation: three short chains, summed in fp32 by the epilogue, drift less
//     than one long one);
//   * the epilogue reads the three accumulators from TMEM, adds bias / applies ReLU and stores, or -- split-K --
//     writes a partial tile, takes a ticket, and the last CTA of the tile sums the partials in split order (deterministic).
#include <cstdlib>
#include "common.cuh"

namespace moma {
namespace gtc {

constexpr int BM = 128, BN = 64, BK = 32, kStages = 3, kProducers = 256, kThreads = 32 + kProducers;
constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;           // A hi (raw), A lo, B hi (raw), B lo
constexpr int kTmemCols = 256;                                     // 3 accumulators x 64 columns
constexpr int kMaxSplits = 16;

__device__ int g_gtc_error = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) { atomicExch(&g_gtc_error, code); __trap(); }
    }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// layout: 2 = 128-byte swizzle (16-byte chunks XOR row % 8; K-major operands), 1 = 128-byte swizzle with a 32-byte base
// (32-byte chunks XOR row % 4) -- the only shared-memory layout the tensor core reads MN-major TF32 operands from
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) { return make_desc(saddr, 16, 1024, 2); }
// MN-major: [block of 32 m / n = one 4 KiB TMA box][groups of 4 k rows][128 B]: LBO = bytes between the 32-wide blocks,
// SBO = bytes between the 4-row k groups
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) { return make_desc(saddr, 4096, 512, 1); }
// kind::tf32: c_format F32 (bit 4), a_format = b_format = 2 (TF32), majors at bits 15 / 16, N >> 3 at 17, M >> 4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
#define GTC_LD32(a, r)                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                               \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                                \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"              \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),   \
                   "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),            \
                   "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),         \
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),         \
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),         \
                   "=r"(r[31])                                                                           \
                 : "r"(a) : "memory")
#define GTC_WAIT_LD32(r)                                                                                 \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                        \
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),   \
                   "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]),            \
                   "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),         \
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),         \
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),         \
                   "+r"(r[31]) :: "memory")

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    const uint32_t h0 = (__float_as_uint(v.x) + 0x1000u) & 0xffffe000u, h1 = (__float_as_uint(v.y) + 0x1000u) & 0xffffe000u;
    const uint32_t h2 = (__float_as_uint(v.z) + 0x1000u) & 0xffffe000u, h3 = (__float_as_uint(v.w) + 0x1000u) & 0xffffe000u;
    hi = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(h2), __uint_as_float(h3));
    lo = make_float4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
}

struct __align__(16) Bars {
    uint64_t split_done[kStages];
    uint64_t stage_free[kStages];
    uint64_t acc_full;
    uint32_t tmem_base;
    uint32_t last;
};
constexpr int SMEM_TOTAL = kStages * STAGE_BYTES + 1024 + (int)sizeof(Bars);

struct Params {
    const float* A; const float* B;
    long long lda, ldb;
    const float* bias;         // [N] or null
    const float* mask;         // optional, A's stored shape: A elements with mask <= 0 read as 0 (ReLU backward)
    float* C;
    long long ldc, ldm;
    int M, N, K, relu, splits;
    float* partial;            // [splits, tiles, BN, BM] when splits > 1
    unsigned int* tickets;     // one per output tile, zero between launches
};

// One operand tile of a K-slab in registers: R float4 per producer thread (+ the mask of A), and where they go.
// ROWS = 128 (A) or 64 (B).  K-major storage [rows][K]: chunk = 16 bytes of one row's k; MN-major storage [K][rows]:
// chunk = 16 bytes (4 consecutive rows) of one k.  Everything but the slab's k offset is fixed per thread:
//   goff  element offset of the chunk in slab 0          kstep  element offset between consecutive slabs
//   soff  byte offset inside the shared-memory plane      klim   the chunk is inside the matrix while k0 < klim
template <int ROWS, int MN>
struct TileIo {
    static constexpr int R = ROWS * BK / 4 / kProducers;          // float4 per thread: 4 (A), 2 (B)
    __device__ static __forceinline__ void locate(int tid, int j, int r0, int rows, int K, long long ld, long long& goff, int& soff, int& klim) {
        const int idx = tid + j * kProducers;
        if (MN == 0) {
            const int row = idx >> 3, ch = idx & 7;
            goff = (long long)(r0 + row) * ld + 4 * ch;
            soff = row * 128 + ((ch ^ (row & 7)) << 4);                               // 128-byte swizzle
            klim = (r0 + row < rows) ? K - 4 * ch : 0;
        } else {
            constexpr int CPR = ROWS / 4;                                            // chunks per k row: 32 (A), 16 (B)
            const int krow = idx / CPR, ch = idx % CPR;
            goff = (long long)krow * ld + r0 + 4 * ch;
            soff = (ch >> 3) * 4096 + krow * 128 + (((ch & 7) ^ ((krow & 3) << 1)) << 4);   // 128-byte swizzle, 32-byte base
            klim = (r0 + 4 * ch < rows) ? K - krow : 0;
        }
    }
    __device__ static __forceinline__ long long kstep(long long ld) { return MN == 0 ? (long long)BK : (long long)BK * ld; }
};

// A_MN: 0 = A is [M, K] with K contiguous, 1 = A is stored [K, M] with M contiguous (d weight: A = dY^T)
// B_MN: 0 = B is [N, K] with K contiguous (forward), 1 = B is [K, N] with N contiguous (d input, d weight)
template <int A_MN, int B_MN, bool MASK>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN, split = blockIdx.z;
    const int kslabs = (p.K + BK - 1) / BK;
    const int s0 = (int)((long long)kslabs * split / p.splits), s1 = (int)((long long)kslabs * (split + 1) / p.splits);
    const int ns = s1 - s0;

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    Bars* bars = reinterpret_cast<Bars*>(base_ptr + kStages * STAGE_BYTES);
    auto a_hi = [&](int s) { return base + s * STAGE_BYTES; };
    auto a_lo = [&](int s) { return base + s * STAGE_BYTES + A_BYTES; };
    auto b_hi = [&](int s) { return base + s * STAGE_BYTES + 2 * A_BYTES; };
    auto b_lo = [&](int s) { return base + s * STAGE_BYTES + 2 * A_BYTES + B_BYTES; };
    auto bar_split = [&](int s) { return smem_u32(&bars->split_done[s]); };
    auto bar_free = [&](int s) { return smem_u32(&bars->stage_free[s]); };
    const uint32_t bar_acc = smem_u32(&bars->acc_full);

    if (warp == 0) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) { mbar_init(bar_split(s), kProducers / 32); mbar_init(bar_free(s), 1); }
            mbar_init(bar_acc, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = bars->tmem_base;
    pdl_wait();

    if (warp == 0) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            // the hi and lo planes of B are adjacent in the stage: ONE instruction with N = 2 BN multiplies A_hi by both
            // (A_hi is read from shared memory once for hi.hi and hi.lo)
            constexpr uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN), idesc2 = make_idesc(BM, 2 * BN, A_MN, B_MN);
            for (int i = 0; i < ns; ++i) {
                const int s = i % kStages;
                mbar_wait(bar_split(s), (i / kStages) & 1, 602);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t dah = A_MN ? desc_mnmajor(a_hi(s)) : desc_kmajor(a_hi(s));
                const uint64_t dal = A_MN ? desc_mnmajor(a_lo(s)) : desc_kmajor(a_lo(s));
                const uint64_t dbh = B_MN ? desc_mnmajor(b_hi(s)) : desc_kmajor(b_hi(s));
            #pragma unroll
                for (int ks = 0; ks < BK / 8; ++ks) {
                    const uint64_t kk = (uint64_t)((ks * 32) >> 4);                          // 8 floats along the 128-byte row
                    const uint64_t kg = (uint64_t)((ks * 1024) >> 4);                        // next 8 k rows
                    const uint64_t ao = A_MN ? kg : kk, bo = B_MN ? kg : kk;
                    const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                    umma_tf32(tmem + 0, dal + ao, dbh + bo, idesc, acc);                     // lo . hi        -> columns [0, 64)
                    umma_tf32(tmem + 64, dah + ao, dbh + bo, idesc2, acc);                   // hi . (hi | lo) -> columns [64, 192)
                }
                umma_commit(bar_free(s));
            }
            umma_commit(bar_acc);
        }
    } else {
        // ===================================================== producers (warps 1-8): global -> registers -> hi / lo planes
        const int tid = threadIdx.x - 32;                          // 0..255
        using TA = TileIo<BM, A_MN>;
        using TB = TileIo<BN, B_MN>;
        // three register buffers: slab i is split and stored while slabs i + 1 and i + 2 are in flight from global memory
        float4 a0[TA::R], a1[TA::R], a2[TA::R], b0[TB::R], b1[TB::R], b2[TB::R];
        float4 q0[MASK ? TA::R : 1], q1[MASK ? TA::R : 1], q2[MASK ? TA::R : 1];        // ReLU masks of A
        int sa[TA::R], sb[TB::R], la[TA::R], lb[TB::R];            // shared-memory offsets, k limits (same every slab)
        const float *pa[TA::R], *pq[MASK ? TA::R : 1], *pb[TB::R]; // chunk addresses in slab 0
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f), one = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
        for (int j = 0; j < TA::R; ++j) {
            long long g;
            TA::locate(tid, j, m0, p.M, p.K, p.lda, g, sa[j], la[j]);
            pa[j] = p.A + g;
            if (MASK) { int d0, d1; TA::locate(tid, j, m0, p.M, p.K, p.ldm, g, d0, d1); pq[j] = p.mask + g; }
        }
#pragma unroll
        for (int j = 0; j < TB::R; ++j) {
            long long g;
            TB::locate(tid, j, n0, p.N, p.K, p.ldb, g, sb[j], lb[j]);
            pb[j] = p.B + g;
        }
        const long long ka = TA::kstep(p.lda), kq = TA::kstep(p.ldm), kb = TB::kstep(p.ldb);
        auto fetch = [&](int slab, float4* ra, float4* rm, float4* rb) {
            const int k0 = slab * BK;
#pragma unroll
            for (int j = 0; j < TA::R; ++j) {
                const bool ok = k0 < la[j];
                ra[j] = ok ? __ldg(reinterpret_cast<const float4*>(pa[j] + slab * ka)) : zero;
                if (MASK) rm[j] = ok ? __ldg(reinterpret_cast<const float4*>(pq[j] + slab * kq)) : one;
            }
#pragma unroll
            for (int j = 0; j < TB::R; ++j)
                rb[j] = (k0 < lb[j]) ? __ldg(reinterpret_cast<const float4*>(pb[j] + slab * kb)) : zero;
        };
        // slab i: start the loads of slab i + 2 into the buffer slab i - 1 left, then split / store this one
        auto step = [&](int i, const float4* ca, const float4* cm, const float4* cb, float4* na, float4* nm, float4* nb) {
            if (i + 2 < ns) fetch(s0 + i + 2, na, nm, nb);
            const int s = i % kStages;
            mbar_wait(bar_free(s), ((i / kStages) & 1) ^ 1, 601);
            uint8_t* st = base_ptr + s * STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < TA::R; ++j) {
                float4 v = ca[j];
                if (MASK) {
                    v.x = cm[j].x > 0.f ? v.x : 0.f; v.y = cm[j].y > 0.f ? v.y : 0.f;
                    v.z = cm[j].z > 0.f ? v.z : 0.f; v.w = cm[j].w > 0.f ? v.w : 0.f;
                }
                float4 hi, lo;
                split4(v, hi, lo);
                *reinterpret_cast<float4*>(st + sa[j]) = hi;
                *reinterpret_cast<float4*>(st + A_BYTES + sa[j]) = lo;
            }
#pragma unroll
            for (int j = 0; j < TB::R; ++j) {
                float4 hi, lo;
                split4(cb[j], hi, lo);
                *reinterpret_cast<float4*>(st + 2 * A_BYTES + sb[j]) = hi;
                *reinterpret_cast<float4*>(st + 2 * A_BYTES + B_BYTES + sb[j]) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic writes -> visible to the MMA (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_split(s));
        };
        if (ns > 0) fetch(s0, a0, q0, b0);
        if (ns > 1) fetch(s0 + 1, a1, q1, b1);
        for (int i = 0; i < ns; i += 3) {
            step(i, a0, q0, b0, a2, q2, b2);
            if (i + 1 < ns) step(i + 1, a1, q1, b1, a0, q0, b0);
            if (i + 2 < ns) step(i + 2, a2, q2, b2, a1, q1, b1);
        }
        pdl_launch_dependents();
        // ---- epilogue: thread = one row x 32 columns of the 128 x 64 tile (TMEM lane quarter = warp % 4)
        mbar_wait(bar_acc, 0, 604);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int half = (warp - 1) >> 2;                          // warps 1-4: columns 0-31, warps 5-8: columns 32-63
        const int rit = (warp & 3) * 32 + lane;
        const int row = m0 + rit, c0 = n0 + 32 * half;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        float out[32];
        {
            uint32_t r0[32], r1[32], r2[32];
            GTC_LD32(tmem + lane_off + 0 + 32 * half, r0);                       // lo . hi
            GTC_LD32(tmem + lane_off + 128 + 32 * half, r1);                     // hi . lo
            GTC_LD32(tmem + lane_off + 64 + 32 * half, r2);                      // hi . hi
            GTC_WAIT_LD32(r0); GTC_WAIT_LD32(r1); GTC_WAIT_LD32(r2);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                out[j] = (__uint_as_float(r0[j]) + __uint_as_float(r1[j])) + __uint_as_float(r2[j]);
        }
        const bool vec_ok = (p.ldc % 4 == 0) && (c0 + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
        auto finish = [&]() {                                      // bias / ReLU / store of this thread's 32 columns
            if (row >= p.M) return;
            float* crow = p.C + (long long)row * p.ldc + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float v = out[j] + ((p.bias != nullptr && c0 + j < p.N) ? __ldg(p.bias + c0 + j) : 0.f);
                out[j] = p.relu ? fmaxf(v, 0.f) : v;
            }
            if (vec_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(crow + j) = make_float4(out[j], out[j + 1], out[j + 2], out[j + 3]);
            } else {
                for (int j = 0; j < 32; ++j) if (c0 + j < p.N) crow[j] = out[j];
            }
        };
        if (p.splits == 1) {
            finish();
        } else {
            // split-K: partial tile -> workspace [split][tile][column][row in tile] (lanes = consecutive rows: coalesced),
            // ticket, the last CTA of the tile sums the partials in split order (deterministic)
            const int tile = blockIdx.y * gridDim.x + blockIdx.x;
            const long long tiles = (long long)gridDim.x * gridDim.y;
            float* pt = p.partial + ((long long)split * tiles + tile) * (BM * BN) + (32 * half) * BM + rit;
#pragma unroll
            for (int j = 0; j < 32; ++j) pt[j * BM] = out[j];
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid == 0) bars->last = (atomicAdd(&p.tickets[tile], 1u) == (unsigned)p.splits - 1) ? 1u : 0u;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (bars->last != 0u) {
                __threadfence();
#pragma unroll
                for (int j = 0; j < 32; ++j) out[j] = 0.f;
                for (int sp = 0; sp < p.splits; ++sp) {
                    const float* ps = p.partial + ((long long)sp * tiles + tile) * (BM * BN) + (32 * half) * BM + rit;
#pragma unroll
                    for (int j = 0; j < 32; ++j) out[j] += __ldcg(ps + j * BM);
                }
                finish();
                if (tid == 0) p.tickets[tile] = 0u;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace gtc

// shapes this kernel takes: enough rows for a 128-row accumulator, operands readable as aligned float4
static bool gemm_tc_shape_ok(const float* A, long long lda, const float* B, long long ldb, int M, int N, int K) {
    if (M < 128 || N < 32 || K < 32) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda % 4) || (ldb % 4)) return false;
    return true;
}
// The product path uses this kernel from 2 M N K = 2^28 flop (512 x 512 x 512) up.  Measured on the B200, back-to-back launches in
// a graph, warp-level kernel of gemm.cu -> this one: 512x512x512 12.6 -> 9.8 us, 512x128x2048 18.1 -> 14.6, 512x2048x128 10.5 ->
// 6.8, 1024x512x512 18.8 -> 10.9, 512x2048x2048 110 -> 51; below that the longer hand-off chain of this kernel (global ->
// registers -> shared -> tensor core -> TMEM -> registers) loses to the warp-level kernel's ~1 us body (512x384x128: 4.6 -> 7.3).
// MOMA_B200_GEMM_TC=0 keeps every GEMM on the warp-level kernel, =all sends every supported shape here.
bool gemm_tc_supported(const float* A, long long lda, const float* B, long long ldb, int M, int N, int K) {
    static const int mode = [] { const char* e = getenv("MOMA_B200_GEMM_TC"); return e == nullptr ? 1 : (e[0] == '0' ? 0 : (e[0] == 'a' ? 2 : 1)); }();
    if (mode == 0 || !gemm_tc_shape_ok(A, lda, B, ldb, M, N, K)) return false;
    return mode == 2 || 2.0 * M * N * K >= 268435456.0;
}
int gemm_tc_splits(int M, int N, int K) {
    const int tiles = ((M + gtc::BM - 1) / gtc::BM) * ((N + gtc::BN - 1) / gtc::BN);
    const int kslabs = (K + gtc::BK - 1) / gtc::BK;
    int s = sm_count() / tiles;
    static const int min_slabs = [] { const char* e = getenv("MOMA_B200_GEMM_TC_MINSLABS"); int v = e ? atoi(e) : 4; return v < 1 ? 1 : v; }();
    if (s > kslabs / min_slabs) s = kslabs / min_slabs;   // at least four slabs per split: the fix-up costs ~2.5 us of latency
    // the tensor core accumulates with truncation: keep one accumulation chain to <= 16 slabs (512 k) -- measured error
    // against fp64 at K = 2048 in one chain 5.9e-6 of max|C|, in four chains 4e-7
    const int smin = (kslabs + 15) / 16;
    if (s < smin) s = smin;
    if (s > gtc::kMaxSplits) s = gtc::kMaxSplits;
    return s < 1 ? 1 : s;
}
size_t gemm_tc_workspace_bytes(int M, int N, int K) {
    const int s = gemm_tc_splits(M, N, K);
    const size_t tiles = (size_t)((M + gtc::BM - 1) / gtc::BM) * ((N + gtc::BN - 1) / gtc::BN);
    return ((tiles * sizeof(unsigned) + 255) / 256) * 256 + (s > 1 ? (size_t)s * tiles * gtc::BM * gtc::BN * sizeof(float) : 0);
}
// a_mn = 0: A is [M, K] (lda = row stride, K contiguous); a_mn = 1: A is stored [K, M] (M contiguous)
// b_mn = 0: B is [N, K] (ldb = row stride, K contiguous); b_mn = 1: B is [K, N] (N contiguous)
// mask (nullable) has A's storage shape and row stride ldm
int gemm_tc(const float* A, long long lda, int a_mn, const float* mask, long long ldm, const float* B, long long ldb, int b_mn,
            const float* bias, float* C, long long ldc, int M, int N, int K, int relu, void* workspace, size_t workspace_bytes,
            cudaStream_t st) {
    using namespace gtc;
    note_flops(0, 2.0 * M * N * K);
    // float4 granularity along the contiguous dimension of each operand
    MOMA_REQUIRE((a_mn ? M : K) % 4 == 0 && (b_mn ? N : K) % 4 == 0 && (mask == nullptr || (aligned16(mask) && ldm % 4 == 0)),
                 MOMA_ERR_ALIGN, "gemm_tc: contiguous dimensions must be multiples of 4 floats");
    Params p{};
    p.A = A; p.B = B; p.lda = lda; p.ldb = ldb;
    p.bias = bias; p.mask = mask; p.C = C; p.ldc = ldc; p.ldm = ldm; p.M = M; p.N = N; p.K = K; p.relu = relu;
    p.splits = 1; p.partial = nullptr; p.tickets = nullptr;
    if (workspace != nullptr) {
        const int s = gemm_tc_splits(M, N, K);
        if (s > 1) {
            MOMA_REQUIRE(workspace_bytes >= gemm_tc_workspace_bytes(M, N, K) && aligned16(workspace), MOMA_ERR_WORKSPACE,
                         "gemm_tc: workspace too small or unaligned");
            const size_t tiles = (size_t)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
            p.splits = s;
            p.tickets = static_cast<unsigned int*>(workspace);
            p.partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + ((tiles * sizeof(unsigned) + 255) / 256) * 256);
        }
    }
    const dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, p.splits);
    auto go = [&](auto kernel) {
        ensure_dyn_smem(reinterpret_cast<const void*>(kernel), SMEM_TOTAL);
        launch_pdl(kernel, grid, dim3(kThreads), (size_t)SMEM_TOTAL, st, p);
    };
    if (a_mn && !b_mn) return fail(MOMA_ERR_UNSUPPORTED, "gemm_tc: A^T . B^T is not instantiated");
    if (mask != nullptr) {
        if (a_mn) go(gemm_tc_kernel<1, 1, true>);
        else if (b_mn) go(gemm_tc_kernel<0, 1, true>);
        else go(gemm_tc_kernel<0, 0, true>);
    } else {
        if (a_mn) go(gemm_tc_kernel<1, 1, false>);
        else if (b_mn) go(gemm_tc_kernel<0, 1, false>);
        else go(gemm_tc_kernel<0, 0, false>);
    }
    MOMA_CUDA_LAUNCH_CHECK("gemm_tc");
    return MOMA_OK;                                  // (the callers count the launch, as for gemm_nt)
}

}  // namespace moma

using namespace moma;

// Test entry point: C = act(op(A) . op(B) + bias) through the tcgen05 kernel (a_mn / b_mn: see gemm_tc above).
extern "C" __attribute__((visibility("default"))) int moma_debug_gemm_tc(
    const float* A, int64_t lda, int a_mn, const float* mask, const float* B, int64_t ldb, int b_mn, const float* bias, float* C, int64_t ldc,
    int64_t M, int64_t N, int64_t K, int relu, void* workspace, size_t workspace_bytes, moma_stream_t stream) {
    MOMA_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, MOMA_ERR_INVALID, "debug_gemm_tc: bad arguments");
    MOMA_REQUIRE(gemm_tc_shape_ok(A, lda, B, ldb, (int)M, (int)N, (int)K), MOMA_ERR_UNSUPPORTED, "debug_gemm_tc: unsupported shape");
    return gemm_tc(A, lda, a_mn, mask, lda, B, ldb, b_mn, bias, C, ldc, (int)M, (int)N, (int)K, relu, workspace, workspace_bytes,
                   as_stream(stream));
}
extern "C" __attribute__((visibility("default"))) size_t moma_debug_gemm_tc_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    return gemm_tc_workspace_bytes((int)M, (int)N, (int)K);
}
extern "C" __attribute__((visibility("default"))) int moma_debug_gemm_tc_error(void) {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, gtc::g_gtc_error, sizeof(int)) != cudaSuccess) { cudaGetLastError(); return -1; }
    return v;
}
