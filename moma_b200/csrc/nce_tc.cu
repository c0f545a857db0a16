// InfoNCE logits + cross-entropy forward/backward, BF16 tensor-core path for sm_100a.
//
// One flash-style pass over the (bf16 shadow of the) queue per query tile:
//     S = Q . Tile^T        tcgen05.mma  (A = Q smem, B = queue tile smem, both K-major)  -> TMEM
//     P = exp2(S*c - m)     softmax warps: tcgen05.ld -> registers -> bf16 -> tcgen05.st   -> TMEM
//     O += P . Tile         tcgen05.mma  (A = P from TMEM, B = the SAME smem tile, MN-major) -> TMEM
// so the [B, K] logits never leave the SM and the queue tile is fetched once (TMA, 128B swizzle)
// for both contractions.  The running (m, l, O) per row are written as per-split partials and
// merged (splits x shards, + positive column) by moma_nce_combine in nce_simt.cu.
//
// Replaces MoMA/mem_moco.py:29-49, learning/contrast_trainer.py:189-205 and their backward.
//
// CTA layout (warp-specialised, 1 CTA / SM):
//   warp 0      TMA producer (Q once, queue tiles through a 4-stage mbarrier ring)
//   warp 1      MMA issuer (one elected lane issues every tcgen05.mma / tcgen05.commit)
//   warp 2      TMEM allocator (512 columns)
//   warps 4-7   softmax group 0, warps 8-11 softmax group 1 (one thread per query row)
// (This header describes the first-generation kernel nce_tc_kernel, still used for D = 256; D = 64 / 128 run the
//  v3 kernel further down, which documents its own layout.)
// Two schedules share the code (template NQ):
//   NQ = 2 (D <= 128, B > 128): two 128-row query tiles per CTA ping-pong, so the tensor pipe
//           works on tile B while the softmax group of tile A is busy (TMEM: S0 S1 O0 O1).
//   NQ = 1 (D = 256 or B <= 128): one query tile, S double-buffered, one softmax group.
// Online softmax uses a lazily updated reference max (rescale O only when the row max grows
// by more than 2^8), and tracks the true max separately for the top-1 flag.
#include <cuda.h>
#include <math_constants.h>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include "common.cuh"

namespace moma {
namespace tc {

constexpr int kBM = 128;          // query rows per M-tile (UMMA M)
constexpr int kStages = 4;        // queue-tile ring depth
constexpr int kTmemCols = 512;
constexpr float kLazyTau = 20.0f; // log2 units: P <= 2^20, far inside fp32 / bf16 range; rescales become rare

__device__ int g_tc_error = 0;

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU -- record a code and trap instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            atomicExch(&g_tc_error, code);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

#define TMEM_LD32(a, r)                                                                                  \
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

#define TMEM_ST32(a, r)                                                                                  \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                         \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"                               \
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"                     \
                 ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),  \
                   "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),      \
                   "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),   \
                   "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),   \
                   "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

#define TMEM_ST16(a, r)                                                                                  \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                         \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                              \
                 ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),  \
                   "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]),      \
                   "r"(r[14]), "r"(r[15]) : "memory")

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);      // .x (low 16 bits) = lo
    return *reinterpret_cast<uint32_t*>(&h);
}

// UMMA shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32
//   [4,6) c_format=1 (F32) | [7,10) a_format=1 (BF16) | [10,13) b_format=1 | bit15 a_major | bit16 b_major
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct __align__(8) Bars {
    uint64_t q_full;
    uint64_t kv_full[kStages];
    uint64_t kv_empty[kStages];
    uint64_t s_full[2];
    uint64_t p_full[2];
    uint64_t pv_done[2];
    uint64_t o_final;      // committed once after the last PV: the epilogue must not reuse the
                           // per-tile pv_done phases (with NQ == 1 a warp can be two phases ahead
                           // of it, and a parity wait would alias)
    uint32_t tmem_base;
    uint32_t pad;
};

template <int D, int NQ, int BN>
struct Cfg {
    static constexpr int KB = D / 64;                   // 64-column (128-byte) blocks per row
    static constexpr int Q_BLOCK = kBM * 128;           // bytes of one 64-col block of a Q tile
    static constexpr int Q_TILE = KB * Q_BLOCK;
    static constexpr int K_BLOCK = BN * 128;
    static constexpr int K_TILE = KB * K_BLOCK;
    static constexpr int SMEM_DATA = NQ * Q_TILE + kStages * K_TILE;
    static constexpr int SMEM_TOTAL = SMEM_DATA + 1024 /*align slack*/ + (int)sizeof(Bars);
    static constexpr int THREADS = 128 * (1 + NQ);
    static constexpr int S_COL0 = 0, S_COL1 = 128, O_COL = 256;   // TMEM columns
    static_assert(256 + NQ * D <= kTmemCols, "TMEM budget");
    static_assert(SMEM_TOTAL <= 227 * 1024, "smem budget");
};

template <int D, int NQ, int BN>
__global__ void __launch_bounds__(128 * (1 + NQ), 1)
nce_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              int B, long long K_local, float scale_log2, int n_splits, float* __restrict__ part_m,
              float* __restrict__ part_l, float* __restrict__ part_mmax, float* __restrict__ part_O,
              float* __restrict__ dbg_S) {
    using C = Cfg<D, NQ, BN>;
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mg = blockIdx.x, split = blockIdx.y;
    const int T_total = (int)((K_local + BN - 1) / BN);
    const int t0 = (int)((long long)T_total * split / n_splits);
    const int t1 = (int)((long long)T_total * (split + 1) / n_splits);
    const int nt = t1 - t0;
    const int row_base = mg * NQ * kBM;
    const int nq_active = (NQ == 2 && row_base + kBM < B) ? 2 : 1;

    if (nt <= 0) {      // empty split: neutral partials
        if (warp >= 4) {
            const int g = (warp - 4) >> 2;
            const int row = row_base + g * kBM + ((warp & 3) << 5) + lane;
            if (row < B) {
                const long long o = (long long)split * B + row;
                part_m[o] = -CUDART_INF_F; part_mmax[o] = -CUDART_INF_F; part_l[o] = 0.f;
                for (int d = 0; d < D; ++d) part_O[o * D + d] = 0.f;
            }
        }
        return;
    }

    // carve shared memory (SWIZZLE_128B tiles need 1024-byte alignment)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t q_smem = base;
    const uint32_t kv_smem = base + NQ * C::Q_TILE;
    Bars* bars = reinterpret_cast<Bars*>(smem_raw + (base - raw) + C::SMEM_DATA);
    const uint32_t b_q_full = smem_u32(&bars->q_full);
    auto b_kv_full = [&](int s) { return smem_u32(&bars->kv_full[s]); };
    auto b_kv_empty = [&](int s) { return smem_u32(&bars->kv_empty[s]); };
    auto b_s_full = [&](int b) { return smem_u32(&bars->s_full[b]); };
    auto b_p_full = [&](int b) { return smem_u32(&bars->p_full[b]); };
    auto b_pv_done = [&](int g) { return smem_u32(&bars->pv_done[g]); };
    const uint32_t b_o_final = smem_u32(&bars->o_final);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        mbar_init(b_q_full, 1);
        for (int s = 0; s < kStages; ++s) { mbar_init(b_kv_full(s), 1); mbar_init(b_kv_empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(b_s_full(b), 1); mbar_init(b_p_full(b), 128); mbar_init(b_pv_done(b), 1); }
        mbar_init(b_o_final, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    auto s_colf = [&](int b) { return tmem + (uint32_t)(b * C::S_COL1); };

    // register budget: the producer/issuer warpgroup gives registers to the softmax warpgroups,
    // which hold a whole 128-wide score row per thread
    if (warp < 4) {
      // the producer / issuer warpgroup hands registers to the softmax warpgroups, which hold a whole
      // 128-wide score row per thread
      asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
      if (warp == 0) {
        // ===================================================== TMA producer
        if (lane == 0) {
            mbar_expect_tx(b_q_full, nq_active * C::Q_TILE);
            for (int g = 0; g < nq_active; ++g)
                for (int kb = 0; kb < C::KB; ++kb)
                    tma_load_2d(q_smem + g * C::Q_TILE + kb * C::Q_BLOCK, &tmap_q, kb * 64, row_base + g * kBM, b_q_full);
            for (int i = 0; i < nt; ++i) {
                const int s = i % kStages;
                mbar_wait(b_kv_empty(s), ((i / kStages) & 1) ^ 1, 101);
                mbar_expect_tx(b_kv_full(s), C::K_TILE);
                for (int kb = 0; kb < C::KB; ++kb)
                    tma_load_2d(kv_smem + s * C::K_TILE + kb * C::K_BLOCK, &tmap_k, kb * 64, (t0 + i) * BN, b_kv_full(s));
            }
        }
      } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc(kBM, BN, 0, 0);     // S = Q . Tile^T
            constexpr uint32_t idesc_o = make_idesc(kBM, D, 0, 1);      // O += P . Tile (B MN-major)
            // descriptors advance by adding (bytes >> 4) to the 14-bit start-address field
            const uint64_t q_desc0 = make_desc(q_smem, 16, 1024);
            const uint64_t k_desc0 = make_desc(kv_smem, 16, 1024);                 // K-major view (S)
            const uint64_t v_desc0 = make_desc(kv_smem, C::K_BLOCK, 1024);         // MN-major view (PV)
            auto issue_s = [&](int g_q, int b, int stage) {
                const uint64_t da0 = q_desc0 + (uint64_t)((g_q * C::Q_TILE) >> 4);
                const uint64_t db0 = k_desc0 + (uint64_t)((stage * C::K_TILE) >> 4);
#pragma unroll 1
                for (int ks = 0; ks < D / 16; ++ks) {
                    // 64-column block (ks >> 2), then the 16-element k-step inside the 128-byte swizzle row
                    const uint32_t qoff = (uint32_t)((ks >> 2) * C::Q_BLOCK + (ks & 3) * 32) >> 4;
                    const uint32_t koff = (uint32_t)((ks >> 2) * C::K_BLOCK + (ks & 3) * 32) >> 4;
                    umma_ss(s_colf(b), da0 + qoff, db0 + koff, idesc_s, ks > 0 ? 1u : 0u);
                }
                umma_commit(b_s_full(b));
            };
            auto issue_pv = [&](int g_o, int b, int stage, bool accumulate) {
                const uint64_t db0 = v_desc0 + (uint64_t)((stage * C::K_TILE) >> 4);
                const uint32_t d_tmem = tmem + C::O_COL + g_o * D;
#pragma unroll 1
                for (int ks = 0; ks < BN / 16; ++ks) {
                    // 16 queue rows (= 2 swizzle atoms of 8 rows x 128 B) per k-step
                    umma_ts(d_tmem, s_colf(b) + ks * 8, db0 + (uint64_t)(ks * (2048 >> 4)), idesc_o,
                            (accumulate || ks > 0) ? 1u : 0u);
                }
                umma_commit(b_pv_done(g_o));
            };
            mbar_wait(b_q_full, 0, 102);
            if (NQ == 2) {
                mbar_wait(b_kv_full(0), 0, 103);
                tc_fence_after();
                for (int g = 0; g < nq_active; ++g) issue_s(g, g, 0);
                for (int i = 0; i < nt; ++i) {
                    const int st = i % kStages;
                    for (int g = 0; g < nq_active; ++g) {
                        mbar_wait(b_p_full(g), i & 1, 104);
                        tc_fence_after();
                        issue_pv(g, g, st, i > 0);
                        if (g == nq_active - 1) umma_commit(b_kv_empty(st));
                        if (i + 1 < nt) {
                            const int sn = (i + 1) % kStages;
                            if (g == 0) { mbar_wait(b_kv_full(sn), ((i + 1) / kStages) & 1, 105); tc_fence_after(); }
                            issue_s(g, g, sn);
                        }
                    }
                }
            } else {
                mbar_wait(b_kv_full(0), 0, 103);
                tc_fence_after();
                issue_s(0, 0, 0);
                if (nt > 1) { mbar_wait(b_kv_full(1), 0, 106); tc_fence_after(); issue_s(0, 1, 1); }
                for (int i = 0; i < nt; ++i) {
                    const int st = i % kStages, b = i & 1;
                    mbar_wait(b_p_full(b), (i >> 1) & 1, 104);
                    tc_fence_after();
                    issue_pv(0, b, st, i > 0);
                    umma_commit(b_kv_empty(st));
                    if (i + 2 < nt) {
                        const int sn = (i + 2) % kStages;
                        mbar_wait(b_kv_full(sn), ((i + 2) / kStages) & 1, 105);
                        tc_fence_after();
                        issue_s(0, b, sn);
                    }
                }
            }
            umma_commit(b_o_final);
        }
      }
    } else if ((warp - 4) / 4 < nq_active) {
        // ===================================================== softmax groups
        if (NQ == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int g = (warp - 4) >> 2;
        const int wq = warp & 3;                                 // TMEM lane quarter of this warp
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const int row = row_base + g * kBM + wq * 32 + lane;
        const uint32_t o_addr = tmem + lane_off + C::O_COL + g * D;
        float m_ref = -CUDART_INF_F, m_true = -CUDART_INF_F, l_run = 0.f;
        const bool ragged_last = (K_local % BN) != 0;

        for (int i = 0; i < nt; ++i) {
            const int b = (NQ == 2) ? g : (i & 1);
            const uint32_t ph = (NQ == 2) ? (i & 1) : ((i >> 1) & 1);
            mbar_wait(b_s_full(b), ph, 201);
            tc_fence_after();
            uint32_t v[BN];
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t* vc = v + 32 * c;
                TMEM_LD32(s_colf(b) + lane_off + 32 * c, vc);
            }
            tmem_wait_ld();
            if (dbg_S != nullptr && i == 0 && split == 0 && row < B) {
#pragma unroll
                for (int j = 0; j < BN; ++j) dbg_S[(long long)row * BN + j] = __uint_as_float(v[j]);
            }
            if (ragged_last && (t0 + i) == T_total - 1) {
                const int valid = (int)(K_local - (long long)(T_total - 1) * BN);
#pragma unroll
                for (int j = 0; j < BN; ++j)
                    if (j >= valid) v[j] = __float_as_uint(-CUDART_INF_F);
            }
            float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < BN; j += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(v[j]));     mx1 = fmaxf(mx1, __uint_as_float(v[j + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(v[j + 2])); mx3 = fmaxf(mx3, __uint_as_float(v[j + 3]));
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
            m_true = fmaxf(m_true, mx);
            const bool need = mx > m_ref + kLazyTau;                 // always true on the first tile
#ifdef MOMA_TC_ALWAYS_RESCALE
            if (i > 0) {
#else
            if (i > 0 && __any_sync(0xffffffffu, need)) {
#endif
                if (dbg_S != nullptr && row < B) {
                    float* info = dbg_S + ((long long)B + row) * BN;
                    info[0] += 1.f; info[1] += need ? 1.f : 0.f; info[3] = (float)i;
                    info[2] = need ? ex2(m_ref - mx) : 1.0f;
                }
                // rare: the running max grew by > 2^8 -> rescale O and l to the new reference
                mbar_wait(b_pv_done(g), (i - 1) & 1, 202);
                tc_fence_after();
                const float f = need ? ex2(m_ref - mx) : 1.0f;
#pragma unroll 1
                for (int c = 0; c < D / 32; ++c) {
                    uint32_t o[32];
                    TMEM_LD32(o_addr + 32 * c, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
                    TMEM_ST32(o_addr + 32 * c, o);
                }
                tmem_wait_st();
                l_run *= f;
            }
            if (need) m_ref = mx;
            const float neg = -m_ref;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float p0 = ex2(fmaf(__uint_as_float(v[32 * c + j]), scale_log2, neg));
                    const float p1 = ex2(fmaf(__uint_as_float(v[32 * c + j + 1]), scale_log2, neg));
                    const float p2 = ex2(fmaf(__uint_as_float(v[32 * c + j + 2]), scale_log2, neg));
                    const float p3 = ex2(fmaf(__uint_as_float(v[32 * c + j + 3]), scale_log2, neg));
                    s0 += p0; s1 += p1; s2 += p2; s3 += p3;
                    pk[j >> 1] = pack_bf16(p0, p1);
                    pk[(j >> 1) + 1] = pack_bf16(p2, p3);
                }
                TMEM_ST16(s_colf(b) + lane_off + 16 * c, pk);      // P aliases the S buffer (bf16 pairs)
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(b_p_full(b));
            l_run += (s0 + s1) + (s2 + s3);
        }

        // epilogue: O (TMEM) -> part_O, stats
        mbar_wait(b_o_final, 0, 203);
        tc_fence_after();
        const long long orow = (long long)split * B + row;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            TMEM_LD32(o_addr + 32 * c, o);
            tmem_wait_ld();
            if (row < B) {
                float4* dst = reinterpret_cast<float4*>(part_O + orow * D + 32 * c);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(o[4 * j]), __uint_as_float(o[4 * j + 1]),
                                         __uint_as_float(o[4 * j + 2]), __uint_as_float(o[4 * j + 3]));
            }
        }
        if (row < B) {
            constexpr float kLn2 = 0.6931471805599453f;
            part_m[orow] = m_ref * kLn2;
            part_mmax[orow] = m_true * kLn2;
            part_l[orow] = l_run;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, kTmemCols);
}

// ============================================================================= v3 kernel (D = 64, 128)
// One 128-row query tile per CTA; the score tile S is triple-buffered in TMEM (S0 S1 S2 O = 512 columns at D = 128)
// so the tensor pipe has S(i+1), S(i+2) finished while the softmax of tile i is in flight.  Roles:
//   warp 0 TMA producer (starts loading before the CTA-wide set-up barrier), warp 1 MMA issuer, warp 2 TMEM allocator,
//   warps 4-7   softmax half 0 (score columns  0..63, O columns 0..D/2),
//   warps 8-11  softmax half 1 (score columns 64..127, O columns D/2..D).
// What changed against the round-1 kernel (measured there: 2200 clk per tile, softmax 1940 clk with the MUFU pipe idle
// during its tcgen05.ld / row-max phases, 2.9 us epilogue):
//   * STREAMING softmax: the lazy reference max means P = exp2(S*c - m_ref) does not need this tile's row max, so each
//     32-column chunk goes tcgen05.ld -> scale -> exp2 -> sum -> pack while the next chunk's load is in flight and the
//     other warp of the scheduler keeps the MUFU pipe busy; the row max is only CHECKED at the end of the tile (half-row
//     maxima exchanged through shared memory + a 64-thread named barrier) and the rare tile whose max outgrows the
//     reference by 2^20 (always the first) recomputes its P from the score registers it still holds.
//   * P is written over the half's OWN score columns ([64h, 64h + 32)), so the two halves never wait for each other
//     to have read S; the PV MMAs take their A operand from those two column ranges.
//   * one tcgen05.commit less per tile: PV(i) commits pv_done[i % stages], which both frees the shared-memory stage for
//     the producer and orders O for a rescale (no separate "stage empty" commit).
//   * epilogue through shared memory: O (TMEM) is staged in the (now idle) queue-tile ring with a padded row stride and
//     written out with fully coalesced 128-bit stores (was: one scattered 16-byte store per row per instruction).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d)
        : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)), "l"(*reinterpret_cast<uint64_t*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d)
        : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
// 2^x on the FMA pipe: x = n + f, f in [-0.5, 0.5], cubic minimax for 2^f (max rel. error 7.5e-5, far
// below the bf16 rounding of P), exponent added as an integer.
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float magic = 12582912.0f;                      // 1.5 * 2^23: rounds x to the nearest integer
    const float t = x + magic;
    const float f = x - (t - magic);
    float p = fmaf(f, 0.0551716685f, 0.2426111251f);
    p = fmaf(p, f, 0.6932609677f);
    p = fmaf(p, f, 0.9999280572f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));   // (t - magic) << 23: magic's bits shift out
}
__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void group_barrier(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// tcgen05.wait::ld that also "produces" the registers of the load it waits for, so the compiler cannot move their
// first use above the wait (the load's own asm statement only says the registers are written)
#define TMEM_WAIT_LD32(r)                                                                                \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                        \
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),   \
                   "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]),            \
                   "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),         \
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),         \
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),         \
                   "+r"(r[31]) :: "memory")

#ifdef MOMA_TC_ABLATE
constexpr bool kAblateHooks = true;
#else
constexpr bool kAblateHooks = false;
#endif
constexpr int kStages3 = 5;     // queue-tile ring depth
constexpr int kSBuf = 3;        // score buffers in TMEM: S(i+1), S(i+2) are ready while softmax works on S(i)

struct __align__(16) Bars3 {
    uint64_t q_full;
    uint64_t kv_full[kStages3];
    uint64_t s_full[kSBuf];
    uint64_t p_full[kSBuf];
    uint64_t pv_done[kStages3];  // PV(i) commits pv_done[i % kStages3]; its next completion needs tile i + kStages3 loaded,
                                 // i.e. the producer to have seen this one: a waiter is never more than one phase behind
    uint64_t o_final;
    uint32_t tmem_base;
    uint32_t pad;
    float xch[2][2][kBM];        // [tile parity][half][row] half-row maxima
    float lsum[kBM];             // half-1 row sums for the epilogue
};

template <int D>
struct Cfg3 {
    static constexpr int BN = 128;
    static constexpr int KB = D / 64;
    static constexpr int Q_BLOCK = kBM * 128;
    static constexpr int Q_TILE = KB * Q_BLOCK;
    static constexpr int K_BLOCK = BN * 128;
    static constexpr int K_TILE = KB * K_BLOCK;
    static constexpr int SMEM_DATA = Q_TILE + kStages3 * K_TILE;
    static constexpr int SMEM_TOTAL = SMEM_DATA + 1024 + (int)sizeof(Bars3);
    static constexpr int THREADS = 384;
    static constexpr int O_COL = 128 * kSBuf;
    static constexpr int OH = D / 2;          // O columns per softmax half
    static constexpr int STAGE_LD = D + 4;    // padded row stride (floats) of the epilogue staging: conflict-free float4 rows
    static_assert(128 * kSBuf + D <= kTmemCols && SMEM_TOTAL <= 227 * 1024, "budget");
    static_assert(kBM * STAGE_LD * 4 <= kStages3 * K_TILE, "epilogue staging must fit in the queue-tile ring");
};

// scale -> exp2 -> sum -> pack of one 32-column chunk against the reference max -m = ng2
template <bool kPoly>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&v)[32], const float2 sc2, const float2 ng2, float2& la,
                                              float2& lb, uint32_t (&pk)[16]) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float2 x01 = ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sc2, ng2);
        const float2 x23 = ffma2(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), sc2, ng2);
        const float2 p01 = make_float2(ex2(x01.x), ex2(x01.y));
        const float2 p23 = make_float2(ex2(x23.x), kPoly ? ex2_poly(x23.y) : ex2(x23.y));
        la = fadd2(la, p01);
        lb = fadd2(lb, p23);
        pk[j >> 1] = pack_bf16(p01.x, p01.y);
        pk[(j >> 1) + 1] = pack_bf16(p23.x, p23.y);
    }
}
__device__ __forceinline__ float chunk_max(const uint32_t (&v)[32]) {
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        m0 = fmaxf(fmaxf(m0, __uint_as_float(v[j])), __uint_as_float(v[j + 1]));
        m1 = fmaxf(fmaxf(m1, __uint_as_float(v[j + 2])), __uint_as_float(v[j + 3]));
    }
    return fmaxf(m0, m1);
}

// ---- in-kernel combine (one launch for the whole InfoNCE pass) ---------------------------------------------------------
// After a CTA has written its split partial it takes a ticket on its query tile; when all n_splits CTAs of the tile have
// arrived, EVERY one of them combines a slice of the tile's 128 rows (rows split, split + n_splits, ...): merge of the
// split partials in split order (deterministic), then -- mode kFuseFinal -- the positive column, loss row, d loss / d q and
// the top-1 flag, or -- mode kFusePacked -- one packed record per row for the cross-rank exchange.  The last CTA of the
// GRID (second ticket) reduces the loss rows to the mean / accuracy and zeroes the counters for the next launch.
// Waiting on other CTAs is safe here: the grid never exceeds one CTA per SM (nce_tc_num_splits), so all of its CTAs are
// co-resident as soon as whatever else runs on the GPU has drained; the wait is bounded and traps instead of hanging.
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float bf16_rne(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// `wpr` warps combine one row: warp `sub` of them takes the splits sub, sub + wpr, ... (8 independent 128-bit loads in
// flight per lane and batch), the partial sums meet in shared memory and warp 0 of the row finishes in a fixed order
// (deterministic).  Lane `lane` owns columns [4 * lane, 4 * lane + 4).  `red`: [8 warps][33] float4 scratch.
template <int D>
__device__ __forceinline__ void fuse_combine_row(const NceFuse& f, int row, int B, int n_splits, const float* part_m,
                                                 const float* part_l, const float* part_mmax, const float* part_O, int lane,
                                                 int sub, int wpr, int slot0, float4* red, int bar_id) {
    constexpr int kMaxPer = 5;                         // n_splits <= 160
    float mref = -CUDART_INF_F, mtrue = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < kMaxPer; ++i) {
        const int sidx = lane + 32 * i;
        if (sidx < n_splits) {
            mref = fmaxf(mref, __ldcg(part_m + (long long)sidx * B + row));
            mtrue = fmaxf(mtrue, __ldcg(part_mmax + (long long)sidx * B + row));
        }
    }
    mref = warp_max(mref); mtrue = warp_max(mtrue);
    const bool fin = f.mode == kFuseFinal;
    const bool col_ok = 4 * lane < D;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f);
    float pos = 0.f;
    if (fin) {
        float dot = 0.f;
        if (col_ok) {
            float4 qv = *reinterpret_cast<const float4*>(f.q_f32 + (long long)row * D + 4 * lane);
            kv = *reinterpret_cast<const float4*>(f.kpos_f32 + (long long)row * D + 4 * lane);
            if (f.round_bf16) {
                qv = make_float4(bf16_rne(qv.x), bf16_rne(qv.y), bf16_rne(qv.z), bf16_rne(qv.w));
                kv = make_float4(bf16_rne(kv.x), bf16_rne(kv.y), bf16_rne(kv.z), bf16_rne(kv.w));
            }
            dot = qv.x * kv.x + qv.y * kv.y + qv.z * kv.z + qv.w * kv.w;
        }
        pos = warp_sum(dot) * f.inv_T;
    }
    const float mstar = fin ? fmaxf(mref, pos) : mref;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float lsum = 0.f;
    const float* Op = part_O + (long long)row * D + 4 * lane;
    for (int s0 = sub; s0 < n_splits; s0 += 8 * wpr) {
        float4 v[8];
        float ms[8], ls[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int sidx = s0 + u * wpr;
            const bool ok = sidx < n_splits;
            ms[u] = ok ? __ldcg(part_m + (long long)sidx * B + row) : -CUDART_INF_F;
            ls[u] = ok ? __ldcg(part_l + (long long)sidx * B + row) : 0.f;
            v[u] = (ok && col_ok) ? __ldcg(reinterpret_cast<const float4*>(Op + (long long)sidx * B * D))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float w = ms[u] > -CUDART_INF_F ? expf(ms[u] - mstar) : 0.f;
            lsum = fmaf(w, ls[u], lsum);
            acc.x = fmaf(w, v[u].x, acc.x); acc.y = fmaf(w, v[u].y, acc.y);
            acc.z = fmaf(w, v[u].z, acc.z); acc.w = fmaf(w, v[u].w, acc.w);
        }
    }
    // meet in shared memory: slot0 + sub is this warp's scratch row; element 32 carries the row-sum share
    float4* mine = red + (slot0 + sub) * 33;
    mine[lane] = acc;
    if (lane == 0) mine[32] = make_float4(lsum, 0.f, 0.f, 0.f);
    if (wpr > 1) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 * wpr) : "memory");
    else __syncwarp();
    if (sub != 0) return;
    float l_tot = 0.f;
    acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < wpr; ++j) {
        const float4 o = red[(slot0 + j) * 33 + lane];
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        l_tot += red[(slot0 + j) * 33 + 32].x;
    }
    if (fin) {
        const float wpos = expf(pos - mstar);
        l_tot += wpos;
        if (col_ok) {
            const float sc = f.inv_T * f.dq_scale, il = 1.0f / l_tot;
            float4 o;
            o.x = ((acc.x + wpos * kv.x) * il - kv.x) * sc; o.y = ((acc.y + wpos * kv.y) * il - kv.y) * sc;
            o.z = ((acc.z + wpos * kv.z) * il - kv.z) * sc; o.w = ((acc.w + wpos * kv.w) * il - kv.w) * sc;
            *reinterpret_cast<float4*>(f.dq + (long long)row * D + 4 * lane) = o;
        }
        if (lane == 0) {
            f.loss_rows[row] = logf(l_tot) + mstar - pos;
            f.pos_is_max[row] = pos >= mtrue ? 1 : 0;
            if (f.max_logit) f.max_logit[row] = fmaxf(pos, mtrue);
        }
    } else {
        float* rec = f.packed + (long long)row * (D + 4);
        if (col_ok) *reinterpret_cast<float4*>(rec + 4 * lane) = acc;
        if (lane == 0) *reinterpret_cast<float4*>(rec + D) = make_float4(mref, l_tot, mtrue, 0.f);
    }
}

template <int D, bool kPoly, bool kRagged>
__global__ void __launch_bounds__(384, 1)
nce_tc3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               int B, long long K_local, float scale_log2, int n_splits, float* __restrict__ part_m,
               float* __restrict__ part_l, float* __restrict__ part_mmax, float* __restrict__ part_O,
               float* __restrict__ dbg_S, int hook_arg, const NceFuse fuse) {
    using C = Cfg3<D>;
    constexpr int BN = C::BN;
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x, split = blockIdx.y;
    const int T_total = (int)((K_local + BN - 1) / BN);
    const int t0 = (int)((long long)T_total * split / n_splits);
    const int t1 = (int)((long long)T_total * (split + 1) / n_splits);
    const int nt = t1 - t0;
    const int row_base = mt * kBM;
    // life-cycle trace (hooked build, bit 1024 of MOMA_TC_ABLATE): clock64 stamps of the first and the last CTA of the grid
    long long* life = nullptr;
    if (kAblateHooks && (hook_arg & 1024) && dbg_S != nullptr) {
        const bool first = mt == 0 && split == 0, last = mt == (int)gridDim.x - 1 && split == (int)gridDim.y - 1;
        if (first || last) life = reinterpret_cast<long long*>(dbg_S + 2ll * B * 128) + 4096 + (last ? 32 : 0);
    }
    float* dump_S = (kAblateHooks && (hook_arg & 1024)) ? nullptr : dbg_S;      // the score dump would distort the trace
#define LIFE(k) do { if (kAblateHooks && life) life[k] = clock64(); } while (0)
    if (kAblateHooks && life && threadIdx.x == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        life[0] = clock64(); life[20] = (long long)gt;
    }

    if (nt <= 0) {      // empty split: neutral partials
        pdl_wait();
        if (warp >= 4 && warp < 8) {
            const int row = row_base + ((warp & 3) << 5) + lane;
            if (row < B) {
                const long long o = (long long)split * B + row;
                part_m[o] = -CUDART_INF_F; part_mmax[o] = -CUDART_INF_F; part_l[o] = 0.f;
                for (int d = 0; d < D; ++d) part_O[o * D + d] = 0.f;
            }
        }
        return;
    }

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t q_smem = base;
    const uint32_t kv_smem = base + C::Q_TILE;
    Bars3* bars = reinterpret_cast<Bars3*>(smem_raw + (base - raw) + C::SMEM_DATA);
    const uint32_t b_q_full = smem_u32(&bars->q_full);
    auto b_kv_full = [&](int s) { return smem_u32(&bars->kv_full[s]); };
    auto b_s_full = [&](int b) { return smem_u32(&bars->s_full[b]); };
    auto b_p_full = [&](int b) { return smem_u32(&bars->p_full[b]); };
    auto b_pv_done = [&](int s) { return smem_u32(&bars->pv_done[s]); };
    const uint32_t b_o_final = smem_u32(&bars->o_final);
    auto load_tile = [&](int i) {        // producer lane: queue tile t0 + i -> stage i % kStages3
        const int s = i % kStages3;
        mbar_expect_tx(b_kv_full(s), C::K_TILE);
        for (int kb = 0; kb < C::KB; ++kb)
            tma_load_2d(kv_smem + s * C::K_TILE + kb * C::K_BLOCK, &tmap_k, kb * 64, (t0 + i) * BN, b_kv_full(s));
    };
    const int n_early = nt < kStages3 ? nt : kStages3;      // tiles whose stage is free from the start

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        mbar_init(b_q_full, 1);
        for (int s = 0; s < kStages3; ++s) { mbar_init(b_kv_full(s), 1); mbar_init(b_pv_done(s), 1); }
        for (int b = 0; b < kSBuf; ++b) { mbar_init(b_s_full(b), 1); mbar_init(b_p_full(b), 8); }
        mbar_init(b_o_final, 1);
        fence_barrier_init();
        // the loads do not need TMEM or the other warps: start them before the CTA-wide set-up barrier.  (The barriers
        // they signal are initialised; nobody else touches a barrier before the __syncthreads below.)
        pdl_wait();                       // Q is the preceding kernel's output
        mbar_expect_tx(b_q_full, C::Q_TILE);
        for (int kb = 0; kb < C::KB; ++kb)
            tma_load_2d(q_smem + kb * C::Q_BLOCK, &tmap_q, kb * 64, row_base, b_q_full);
        for (int i = 0; i < n_early; ++i) load_tile(i);
        LIFE(3);
    }
    if (warp == 2) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    auto s_colf = [&](int b) { return tmem + (uint32_t)(b * 128); };
    if (threadIdx.x == 0) LIFE(1);
    pdl_wait();                           // every thread that touches global memory is ordered after the preceding grid
    if (threadIdx.x == 0) LIFE(2);

    if (warp == 0) {
        // ===================================================== TMA producer (remaining tiles)
        if (lane == 0) {
            for (int i = n_early; i < nt; ++i) {
                // stage i % kStages3 is free once PV(i - kStages3) has completed
                mbar_wait(b_pv_done(i % kStages3), ((i / kStages3) & 1) ^ 1, 301);
                load_tile(i);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc(kBM, BN, 0, 0);
            constexpr uint32_t idesc_o = make_idesc(kBM, D, 0, 1);
            const uint64_t q_desc0 = make_desc(q_smem, 16, 1024);
            const uint64_t k_desc0 = make_desc(kv_smem, 16, 1024);
            const uint64_t v_desc0 = make_desc(kv_smem, C::K_BLOCK, 1024);
            auto issue_s = [&](int b, int stage) {
                const uint64_t db0 = k_desc0 + (uint64_t)((stage * C::K_TILE) >> 4);
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint32_t qoff = (uint32_t)((ks >> 2) * C::Q_BLOCK + (ks & 3) * 32) >> 4;
                    const uint32_t koff = (uint32_t)((ks >> 2) * C::K_BLOCK + (ks & 3) * 32) >> 4;
                    umma_ss(s_colf(b), q_desc0 + qoff, db0 + koff, idesc_s, ks > 0 ? 1u : 0u);
                }
                umma_commit(b_s_full(b));
            };
            auto issue_pv = [&](int b, int stage, bool accumulate) {
                const uint64_t db0 = v_desc0 + (uint64_t)((stage * C::K_TILE) >> 4);
#pragma unroll
                for (int ks = 0; ks < BN / 16; ++ks)      // P of keys [64h, 64h + 64) sits in score columns [64h, 64h + 32)
                    umma_ts(tmem + C::O_COL, s_colf(b) + (uint32_t)((ks >> 2) * 64 + (ks & 3) * 8),
                            db0 + (uint64_t)(ks * (2048 >> 4)), idesc_o, (accumulate || ks > 0) ? 1u : 0u);
                umma_commit(b_pv_done(stage));
            };
            mbar_wait(b_q_full, 0, 302);
            mbar_wait(b_kv_full(0), 0, 303);
            tc_fence_after();
            LIFE(4);
            issue_s(0, 0);
            LIFE(5);
            for (int j = 1; j < kSBuf && j < nt; ++j) { mbar_wait(b_kv_full(j), 0, 306); tc_fence_after(); issue_s(j, j); }
            for (int i = 0; i < nt; ++i) {
                const int st = i % kStages3, b = i % kSBuf;
                mbar_wait(b_p_full(b), (i / kSBuf) & 1, 304);
                tc_fence_after();
                issue_pv(b, st, i > 0);
                if (i + kSBuf < nt) {
                    const int sn = (i + kSBuf) % kStages3;
                    mbar_wait(b_kv_full(sn), ((i + kSBuf) / kStages3) & 1, 305);
                    tc_fence_after();
                    issue_s(b, sn);
                }
            }
            umma_commit(b_o_final);
            LIFE(8);
            if (kAblateHooks && life) life[12] = nt;
        }
    } else if (warp >= 4) {
        // ===================================================== softmax (two column halves)
        const int half = (warp - 4) >> 2;
        const int wq = warp & 3;
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const int rit = wq * 32 + lane;                       // row in tile
        const int row = row_base + rit;
        const uint32_t o_addr = tmem + lane_off + C::O_COL + half * C::OH;
        float m_ref = -CUDART_INF_F, m_true = -CUDART_INF_F;
        float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);     // running half-row sums
        const bool ragged_last = kRagged && (K_local % BN) != 0;     // a separate instantiation: the common one has no masking code
        const float2 sc2 = make_float2(scale_log2, scale_log2);

        for (int i = 0; i < nt; ++i) {
            const int b = i % kSBuf;
            const uint32_t s_addr = s_colf(b) + lane_off + half * 64;      // this half's score columns; P goes over the first 32
            mbar_wait(b_s_full(b), (i / kSBuf) & 1, 401);
            tc_fence_after();
            if (i == 0 && warp == 4 && lane == 0) LIFE(6);
            uint32_t v0[32], v1[32], pk0[16], pk1[16];
            TMEM_LD32(s_addr, v0);
            TMEM_WAIT_LD32(v0);
            TMEM_LD32(s_addr + 32, v1);                       // in flight while chunk 0 is processed
            if (dump_S != nullptr && i == 0 && split == 0 && row < B) {
#pragma unroll
                for (int j = 0; j < 32; ++j) dump_S[(long long)row * BN + half * 64 + j] = __uint_as_float(v0[j]);
            }
            const bool ragged = ragged_last && (t0 + i) == T_total - 1;
            const int valid = ragged ? (int)(K_local - (long long)(T_total - 1) * BN) - half * 64 : 64;
            if (kRagged && ragged) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j >= valid) v0[j] = __float_as_uint(-CUDART_INF_F);
            }
            float2 ng2 = make_float2(-m_ref, -m_ref);
            float2 ta = make_float2(0.f, 0.f), tb = make_float2(0.f, 0.f);
            float mxh = chunk_max(v0);
            if (i > 0) softmax_chunk<kPoly>(v0, sc2, ng2, ta, tb, pk0);      // speculative: against the CURRENT reference max
                                                                           // (tile 0 has none yet: its P is computed below)
            TMEM_WAIT_LD32(v1);
            if (dump_S != nullptr && i == 0 && split == 0 && row < B) {
#pragma unroll
                for (int j = 0; j < 32; ++j) dump_S[(long long)row * BN + half * 64 + 32 + j] = __uint_as_float(v1[j]);
            }
            if (kRagged && ragged) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (32 + j >= valid) v1[j] = __float_as_uint(-CUDART_INF_F);
            }
            mxh = fmaxf(mxh, chunk_max(v1));
            if (i > 0) softmax_chunk<kPoly>(v1, sc2, ng2, ta, tb, pk1);
            // the row max is only CHECKED: exchange the half-row maxima with the thread holding the other 64 columns
            bars->xch[i & 1][half][rit] = mxh;
            pair_barrier(1 + wq);
            const float mx = fmaxf(mxh, bars->xch[i & 1][half ^ 1][rit]) * scale_log2;
            m_true = fmaxf(m_true, mx);
            const bool need = mx > m_ref + kLazyTau;                 // identical in both halves of the row; true on tile 0
            if (i == 0 || __any_sync(0xffffffffu, need)) {
                // rare after the first tile: move the reference, rescale O and l, redo this tile's P from the registers
                if (i > 0) {
                    mbar_wait(b_pv_done((i - 1) % kStages3), ((i - 1) / kStages3) & 1, 402);
                    tc_fence_after();
                    const float f = need ? ex2(m_ref - mx) : 1.0f;
#pragma unroll 1
                    for (int c = 0; c < C::OH / 32; ++c) {
                        uint32_t o[32];
                        TMEM_LD32(o_addr + 32 * c, o);
                        TMEM_WAIT_LD32(o);
#pragma unroll
                        for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
                        TMEM_ST32(o_addr + 32 * c, o);
                    }
                    tmem_wait_st();
                    l2a.x *= f; l2a.y *= f; l2b.x *= f; l2b.y *= f;
                }
                if (need) m_ref = mx;
                ng2 = make_float2(-m_ref, -m_ref);
                ta = make_float2(0.f, 0.f); tb = make_float2(0.f, 0.f);
                softmax_chunk<kPoly>(v0, sc2, ng2, ta, tb, pk0);
                softmax_chunk<kPoly>(v1, sc2, ng2, ta, tb, pk1);
            }
            l2a = fadd2(l2a, ta);
            l2b = fadd2(l2b, tb);
            TMEM_ST16(s_addr, pk0);                           // P (bf16 pairs) over this half's own score columns
            TMEM_ST16(s_addr + 16, pk1);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_p_full(b));          // one arrival per warp: 8 instead of 256 smem atomics
            if (i == 0 && warp == 4 && lane == 0) LIFE(7);
        }

        pdl_launch_dependents();     // the combine kernel may start launching; it still waits for this grid to finish
        // epilogue: O (TMEM) -> shared-memory staging (padded rows) -> coalesced 128-bit stores; stats by half 0
        const float l_half = (l2a.x + l2a.y) + (l2b.x + l2b.y);
        if (half == 1) bars->lsum[rit] = l_half;
        mbar_wait(b_o_final, 0, 403);                         // every MMA (and with them every TMA load) has completed:
        tc_fence_after();                                     // the queue-tile ring is free
        if (warp == 4 && lane == 0) LIFE(9);
        float* stage = reinterpret_cast<float*>(smem_raw + (base - raw) + C::Q_TILE);
#pragma unroll 1
        for (int c = 0; c < C::OH / 32; ++c) {
            uint32_t o[32];
            TMEM_LD32(o_addr + 32 * c, o);
            TMEM_WAIT_LD32(o);
            float4* dst = reinterpret_cast<float4*>(stage + rit * C::STAGE_LD + half * C::OH + 32 * c);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                dst[j] = make_float4(__uint_as_float(o[4 * j]), __uint_as_float(o[4 * j + 1]),
                                     __uint_as_float(o[4 * j + 2]), __uint_as_float(o[4 * j + 3]));
        }
        group_barrier(5, 256);                                // all 8 softmax warps: staging complete (also publishes lsum)
        {
            constexpr int V4_PER_ROW = D / 4;
            const int tid = threadIdx.x - 128;                // 0..255
            float* out_base = part_O + ((long long)split * B + row_base) * D;
#pragma unroll 4
            for (int e = tid; e < kBM * V4_PER_ROW; e += 256) {
                const int r = e / V4_PER_ROW, c4 = e - r * V4_PER_ROW;
                if (row_base + r < B)
                    st_stream(reinterpret_cast<float4*>(out_base + (long long)r * D) + c4,
                              *reinterpret_cast<const float4*>(stage + r * C::STAGE_LD + 4 * c4));
            }
        }
        if (half == 0 && row < B) {
            constexpr float kLn2 = 0.6931471805599453f;
            const long long orow = (long long)split * B + row;
            part_m[orow] = m_ref * kLn2;
            part_mmax[orow] = m_true * kLn2;
            part_l[orow] = l_half + bars->lsum[rit];
        }
        if (warp == 4 && lane == 0) LIFE(10);
        if (fuse.mode != kFuseNone) {
            // ---- in-kernel combine: ticket on this query tile, wait for its other splits, combine a slice of its rows
            const int tid = threadIdx.x - 128;
            __threadfence();                                  // this thread's partial stores are visible device-wide
            group_barrier(5, 256);
            if (tid == 0) {
                atomicAdd(&fuse.counters[mt], 1u);
                const long long t_start = clock64();
                while (ld_acquire_u32(&fuse.counters[mt]) < (unsigned)n_splits) {
                    __nanosleep(40);
                    if (clock64() - t_start > 4000000000ll) { atomicExch(&g_tc_error, 501); __trap(); }
                }
            }
            group_barrier(5, 256);
            __threadfence();
            const int w8 = warp - 4;                          // 0..7
            {
                // rows of this CTA's slice: split, split + n_splits, ... (< 128 and < B); `wpr` warps per row
                int n_rows = 0;
                for (int r = split; r < kBM && row_base + r < B; r += n_splits) ++n_rows;
                float4* red = reinterpret_cast<float4*>(stage);          // the staging area is free again (>= 8 x 33 float4)
                for (int j0 = 0; j0 < n_rows; j0 += 8) {
                    const int g = min(8, n_rows - j0);                   // rows handled in this round
                    const int wpr = g == 1 ? 8 : (g == 2 ? 4 : (g <= 4 ? 2 : 1));
                    const int rj = w8 / wpr, sub = w8 % wpr;
                    if (rj < g)
                        fuse_combine_row<D>(fuse, row_base + split + (j0 + rj) * n_splits, B, n_splits, part_m, part_l,
                                            part_mmax, part_O, lane, sub, wpr, rj * wpr, red, 6 + rj);
                    group_barrier(5, 256);                               // scratch reuse between rounds
                }
            }
            // ---- last CTA of the grid: mean loss / accuracy, counters back to zero
            __threadfence();
            group_barrier(5, 256);
            const unsigned total = gridDim.x * gridDim.y;
            if (tid == 0) bars->pad = atomicAdd(&fuse.counters[gridDim.x], 1u) == total - 1 ? 1u : 0u;
            group_barrier(5, 256);
            if (bars->pad != 0u) {
                __threadfence();
                if (w8 == 0) {
                    if (fuse.mode == kFuseFinal && fuse.loss_mean != nullptr) {
                        float sl = 0.f, sa = 0.f;
                        for (int i = lane; i < B; i += 32) { sl += __ldcg(fuse.loss_rows + i); sa += (float)__ldcg(fuse.pos_is_max + i); }
                        sl = warp_sum(sl); sa = warp_sum(sa);
                        if (lane == 0) { *fuse.loss_mean = sl / (float)B; *fuse.acc_pct = sa * (100.0f / (float)B); }
                    }
                    for (int i = lane; i <= (int)gridDim.x; i += 32) fuse.counters[i] = 0u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, kTmemCols);
    if (kAblateHooks && life && threadIdx.x == 0) {
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        life[11] = clock64(); life[21] = (long long)gt;
    }
#undef LIFE
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// [rows, D] bf16 row-major, box = 64 columns (128 B) x box_rows rows, 128-byte swizzle, zero OOB fill
static int make_map(CUtensorMap* out, const void* ptr, long long rows, int D, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    MOMA_REQUIRE(fn != nullptr, MOMA_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (driver too old?)");
    const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)D * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MOMA_REQUIRE(r == CUDA_SUCCESS, MOMA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MOMA_OK;
}

struct MapKey {
    const void* p; long long rows; int D; int box;
    bool operator==(const MapKey& o) const { return p == o.p && rows == o.rows && D == o.D && box == o.box; }
};
struct MapHash {
    size_t operator()(const MapKey& k) const {
        return std::hash<const void*>()(k.p) ^ (std::hash<long long>()(k.rows) * 31) ^ (size_t)(k.D * 131 + k.box);
    }
};
// descriptor cache (the only retained state: keyed by pointer/shape, no device memory held)
static int cached_map(CUtensorMap* out, const void* ptr, long long rows, int D, int box_rows) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapHash> cache;
    const MapKey key{ptr, rows, D, box_rows};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return MOMA_OK; }
    int rc = make_map(out, ptr, rows, D, box_rows);
    if (rc != MOMA_OK) return rc;
    if (cache.size() > 256) cache.clear();
    cache.emplace(key, *out);
    return MOMA_OK;
}

static bool use_v2(int64_t D) {
    static const bool force_v1 = getenv("MOMA_B200_NCE_V1") != nullptr;
    return !force_v1 && (D == 64 || D == 128);
}
static void pick(int64_t B, int64_t D, int* nq, int* bn) {
    *nq = (use_v2(D) || D == 256 || B <= kBM) ? 1 : 2;
    *bn = (D == 256) ? 64 : 128;
}

static bool use_poly_exp() {     // MOMA_B200_NCE_POLY=1: 25 % of the exp2 on the FMA pipe (A/B switch, read once)
    static const bool on = [] { const char* e = getenv("MOMA_B200_NCE_POLY"); return e != nullptr && e[0] == '1'; }();
    return on;
}

template <int D, bool kPoly, bool kRagged>
static int launch3(const void* q, const void* queue, int64_t B, int64_t K_local, float inv_T, int n_splits,
                   float* pm, float* pl, float* pmm, float* pO, float* dbg, cudaStream_t st, const NceFuse& fuse) {
    using C = Cfg3<D>;
    CUtensorMap mq, mk;
    int rc = cached_map(&mq, q, B, D, kBM);
    if (rc != MOMA_OK) return rc;
    rc = cached_map(&mk, queue, K_local, D, C::BN);
    if (rc != MOMA_OK) return rc;
    {
        const cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(nce_tc3_kernel<D, kPoly, kRagged>), C::SMEM_TOTAL);
        if (e != cudaSuccess) take_launch_error();
        MOMA_REQUIRE(e == cudaSuccess, MOMA_ERR_CUDA, "nce_tc3: smem attribute: %s", cudaGetErrorString(e));
    }
    const dim3 grid((unsigned)((B + kBM - 1) / kBM), (unsigned)n_splits);
    const float scale_log2 = inv_T * 1.4426950408889634f;
    int hooks = 0;
    if (kAblateHooks) { const char* e = getenv("MOMA_TC_ABLATE"); hooks = e ? atoi(e) : 0; }
    launch_pdl(nce_tc3_kernel<D, kPoly, kRagged>, grid, dim3(C::THREADS), (size_t)C::SMEM_TOTAL, st, mq, mk, (int)B, (long long)K_local,
               scale_log2, n_splits, pm, pl, pmm, pO, dbg, hooks, fuse);
    MOMA_CUDA_LAUNCH_CHECK("nce_partial(bf16/tcgen05 v3)");
    note_launches(1);
    return MOMA_OK;
}
template <int D>
static int launch2(const void* q, const void* queue, int64_t B, int64_t K_local, float inv_T, int n_splits,
                   float* pm, float* pl, float* pmm, float* pO, float* dbg, cudaStream_t st, const NceFuse& fuse = NceFuse{}) {
    const bool ragged = (K_local % 128) != 0;
    if (use_poly_exp())
        return ragged ? launch3<D, true, true>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st, fuse)
                      : launch3<D, true, false>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st, fuse);
    return ragged ? launch3<D, false, true>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st, fuse)
                  : launch3<D, false, false>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st, fuse);
}

template <int D, int NQ, int BN>
static int launch(const void* q, const void* queue, int64_t B, int64_t K_local, float inv_T, int n_splits,
                  float* pm, float* pl, float* pmm, float* pO, float* dbg, cudaStream_t st) {
    using C = Cfg<D, NQ, BN>;
    CUtensorMap mq, mk;
    int rc = cached_map(&mq, q, B, D, kBM);
    if (rc != MOMA_OK) return rc;
    rc = cached_map(&mk, queue, K_local, D, BN);
    if (rc != MOMA_OK) return rc;
    {
        const cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void*>(nce_tc_kernel<D, NQ, BN>), C::SMEM_TOTAL);
        if (e != cudaSuccess) take_launch_error();
        MOMA_REQUIRE(e == cudaSuccess, MOMA_ERR_CUDA, "nce_tc: smem attribute: %s", cudaGetErrorString(e));
    }
    const dim3 grid((unsigned)((B + NQ * kBM - 1) / (NQ * kBM)), (unsigned)n_splits);
    const float scale_log2 = inv_T * 1.4426950408889634f;
    nce_tc_kernel<D, NQ, BN><<<grid, C::THREADS, C::SMEM_TOTAL, st>>>(mq, mk, (int)B, (long long)K_local, scale_log2,
                                                                     n_splits, pm, pl, pmm, pO, dbg);
    MOMA_CUDA_LAUNCH_CHECK("nce_partial(bf16/tcgen05)");
    note_launches(1);
    return MOMA_OK;
}

static int dispatch(const void* q, const void* queue, int64_t B, int64_t D, int64_t K_local, float inv_T,
                    int n_splits, float* pm, float* pl, float* pmm, float* pO, float* dbg, cudaStream_t st) {
    MOMA_REQUIRE(inv_T > 0.f, MOMA_ERR_INVALID, "nce_tc: temperature must be positive");
    MOMA_REQUIRE((reinterpret_cast<uintptr_t>(q) & 127) == 0 && (reinterpret_cast<uintptr_t>(queue) & 127) == 0,
                 MOMA_ERR_ALIGN, "nce_tc: q / queue must be 128-byte aligned for TMA");
    int nq, bn;
    pick(B, D, &nq, &bn);
    if (use_v2(D)) {
        if (D == 64) return launch2<64>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st);
        return launch2<128>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st);
    }
    if (D == 64) return nq == 2 ? launch<64, 2, 128>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st)
                                : launch<64, 1, 128>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st);
    if (D == 128) return nq == 2 ? launch<128, 2, 128>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st)
                                 : launch<128, 1, 128>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st);
    if (D == 256) return launch<256, 1, 64>(q, queue, B, K_local, inv_T, n_splits, pm, pl, pmm, pO, dbg, st);
    return fail(MOMA_ERR_UNSUPPORTED, "nce_tc: D=%lld unsupported", (long long)D);
}

}  // namespace tc

bool nce_tc_supported(int64_t B, int64_t D, int64_t K_local) {
    return (D == 64 || D == 128 || D == 256) && B > 0 && K_local > 0 && B < (1ll << 30) && K_local < (1ll << 36);
}

int nce_tc_num_splits(int64_t B, int64_t D, int64_t K_local) {
    int nq, bn;
    tc::pick(B, D, &nq, &bn);
    const int64_t mgroups = (B + nq * tc::kBM - 1) / (nq * tc::kBM);
    const int64_t tiles = (K_local + bn - 1) / bn;
    int64_t s = sm_count() / mgroups;
    // at least two queue tiles per CTA: below that a CTA is all fixed cost and the combine kernel reads twice the partials
    // (C2 step: 166.3 us with 1, 162.8 us with 2, 162.7 us with 4); MOMA_B200_NCE_MIN_TILES overrides
    static const int min_tiles = [] { const char* e = getenv("MOMA_B200_NCE_MIN_TILES"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
    if (s > tiles / min_tiles) s = tiles / min_tiles;
    if (s < 1) s = 1;
    return (int)s;
}


// ---- one-launch InfoNCE (tcgen05 partial pass + in-kernel combine); D in {64, 128}
bool nce_fused_supported(int64_t B, int64_t D, int64_t K_local) {
    if (!(D == 64 || D == 128) || B <= 0 || K_local <= 0) return false;
    const int s = nce_tc_num_splits(B, D, K_local);
    const int64_t mt = (B + tc::kBM - 1) / tc::kBM;
    // every CTA of the grid must be co-resident (they wait for each other) and every split must own at least one tile
    return s <= 160 && mt * s <= sm_count() && s <= (K_local + 127) / 128;
}
size_t nce_fused_workspace_floats(int64_t B, int64_t D, int64_t K_local) {
    const int64_t s = nce_tc_num_splits(B, D, K_local);
    return (size_t)(3 * s * B + s * B * D);
}
int nce_fused_launch(const void* q_bf16, const void* queue_bf16, int64_t B, int64_t D, int64_t K_local, float inv_T,
                     float* workspace, const tc::NceFuse& fuse, cudaStream_t stream) {
    const int s = nce_tc_num_splits(B, D, K_local);
    float* pm = workspace;
    float* pl = pm + (int64_t)s * B;
    float* pmm = pl + (int64_t)s * B;
    float* pO = pmm + (int64_t)s * B;
    MOMA_REQUIRE((reinterpret_cast<uintptr_t>(q_bf16) & 127) == 0 && (reinterpret_cast<uintptr_t>(queue_bf16) & 127) == 0,
                 MOMA_ERR_ALIGN, "nce_fused: q / queue must be 128-byte aligned for TMA");
    MOMA_REQUIRE(aligned16(pO), MOMA_ERR_ALIGN, "nce_fused: workspace must be 16-byte aligned");
    if (D == 64) return tc::launch2<64>(q_bf16, queue_bf16, B, K_local, inv_T, s, pm, pl, pmm, pO, nullptr, stream, fuse);
    return tc::launch2<128>(q_bf16, queue_bf16, B, K_local, inv_T, s, pm, pl, pmm, pO, nullptr, stream, fuse);
}

int nce_tc_partial(const void* q, const void* queue, int64_t B, int64_t D, int64_t K_local, float inv_T,
                   int n_splits, float* part_m, float* part_l, float* part_mmax, float* part_O,
                   cudaStream_t stream) {
    return tc::dispatch(q, queue, B, D, K_local, inv_T, n_splits, part_m, part_l, part_mmax, part_O, nullptr, stream);
}

}  // namespace moma

using namespace moma;

// Debug / test entry points (declared in include/moma_b200.h under "debug").
extern "C" __attribute__((visibility("default"))) int moma_debug_nce_tc(
    const void* q, const void* queue, int64_t B, int64_t D, int64_t K_local, float inv_T, int n_splits,
    float* part_m, float* part_l, float* part_mmax, float* part_O, float* dbg_S, moma_stream_t stream) {
    MOMA_REQUIRE(nce_tc_supported(B, D, K_local), MOMA_ERR_UNSUPPORTED, "debug_nce_tc: unsupported shape");
    return tc::dispatch(q, queue, B, D, K_local, inv_T, n_splits, part_m, part_l, part_mmax, part_O, dbg_S,
                        as_stream(stream));
}

// Synchronising read of the device-side protocol-error flag (0 = none).
extern "C" __attribute__((visibility("default"))) int moma_debug_tc_error(void) {
    int v = 0;
    if (cudaMemcpyFromSymbol(&v, tc::g_tc_error, sizeof(int)) != cudaSuccess) { cudaGetLastError(); return -1; }
    return v;
}

// Launch-overhead probe (bench / profiling only): an otherwise empty kernel with a given dynamic shared-memory size,
// optionally allocating and releasing all 512 TMEM columns -- what a launch of the InfoNCE kernel's shape costs
// before it does any work.
namespace moma { namespace tc {
__global__ void __launch_bounds__(384, 1) probe_kernel(int use_tmem, int* sink) {
    extern __shared__ uint8_t probe_smem[];
    __shared__ uint32_t tmem_slot;
    if (use_tmem) {
        if ((threadIdx.x >> 5) == 2) tmem_alloc(smem_u32(&tmem_slot), kTmemCols);
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if ((threadIdx.x >> 5) == 2) tmem_dealloc(tmem_slot, kTmemCols);
    }
    if (sink != nullptr && threadIdx.x == 0 && probe_smem[0] == 123) sink[0] = 1;
}
} }
extern "C" __attribute__((visibility("default"))) int moma_debug_probe_launch(int ctas, int threads, int smem_bytes, int use_tmem,
                                                                              int pdl, moma_stream_t stream) {
    MOMA_REQUIRE(ctas > 0 && threads > 0 && threads <= 384 && smem_bytes >= 0 && smem_bytes <= 227 * 1024, MOMA_ERR_INVALID,
                 "probe_launch: bad arguments");
    cudaError_t e = cudaFuncSetAttribute(tc::probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64);
    MOMA_REQUIRE(e == cudaSuccess, MOMA_ERR_CUDA, "probe_launch: %s", cudaGetErrorString(e));
    if (pdl) launch_pdl(tc::probe_kernel, dim3(ctas), dim3(threads), (size_t)smem_bytes, as_stream(stream), use_tmem, (int*)nullptr);
    else tc::probe_kernel<<<ctas, threads, smem_bytes, as_stream(stream)>>>(use_tmem, nullptr);
    MOMA_CUDA_LAUNCH_CHECK("probe_launch");
    return MOMA_OK;
}
