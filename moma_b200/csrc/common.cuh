// Shared helpers for the moma_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/moma_b200.h"

namespace moma {

// thread-local error string behind moma_last_error()
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define MOMA_REQUIRE(cond, code, ...)                       \
    do {                                                    \
        if (!(cond)) return ::moma::fail((code), __VA_ARGS__); \
    } while (0)

// Launch errors: launch_pdl() records the first failing cudaLaunchKernelEx of the calling thread (take_launch_error()),
// so helpers that launch several kernels and return void still surface it at the entry point's check.
void note_launch_error(cudaError_t e);
cudaError_t take_launch_error();            // returns and clears the calling thread's first recorded launch error
#define MOMA_CUDA_LAUNCH_CHECK(what)                                                   \
    do {                                                                               \
        cudaError_t _e = ::moma::take_launch_error();                                  \
        const cudaError_t _l = cudaGetLastError();                                     \
        if (_e == cudaSuccess) _e = _l;                                                \
        if (_e != cudaSuccess)                                                         \
            return ::moma::fail(MOMA_ERR_CUDA, "%s: %s", (what), cudaGetErrorString(_e)); \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device), thread-safe.
cudaError_t ensure_dyn_smem(const void* func, int bytes);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline cudaStream_t as_stream(moma_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();     // cached multiprocessor count of the current device (148 on B200)
void note_launches(int n);   // launch counter behind moma_debug_launch_count()
void note_flops(int kind, double flops);   // algorithmic-FLOP counters behind moma_debug_flops(): 0 = Linear GEMMs, 1 = attention core

// gemm.cu: C[M,N] = act(A B^T + bias) on the tensor cores (3xTF32); see the file header for the operand convention
int gemm_splits(int M, int N, int K);
size_t gemm_workspace_bytes(int M, int N, int K);
// tcgen05 3xTF32 GEMM (gemm_tc.cu); gemm_nt dispatches to it
bool gemm_tc_supported(const float* A, long long lda, const float* B, long long ldb, int M, int N, int K);
size_t gemm_tc_workspace_bytes(int M, int N, int K);
int gemm_tc(const float* A, long long lda, int a_mn, const float* mask, long long ldm, const float* B, long long ldb, int b_mn,
            const float* bias, float* C, long long ldc, int M, int N, int K, int relu, void* workspace, size_t workspace_bytes,
            cudaStream_t st);
int gemm_nt(const float* A, const float* a_mask, int64_t a_rs, int64_t a_cs, const float* Bm, int64_t b_rs, int64_t b_cs,
            const float* bias, float* C, int64_t ldc, int M, int N, int K, int relu, void* workspace, size_t workspace_bytes,
            cudaStream_t st, void* c_bf16 = nullptr);     // c_bf16: optional dense [M, N] bf16 copy of C
void colsum_masked(const float* X, const float* mask, int rows, int cols, float* out, cudaStream_t st);
bool use_simt_gemm();        // MOMA_B200_GEMM=simt: IEEE-FP32 CUDA-core GEMMs instead (A/B switch, read once)

// Arguments of the in-kernel combine of the tcgen05 InfoNCE kernel (nce_tc.cu; filled by the entry points in nce_simt.cu)
namespace tc {
enum : int { kFuseNone = 0, kFuseFinal = 1, kFusePacked = 2 };
struct NceFuse {
    int mode;
    const float* q_f32;          // [B, D] fp32 queries / positive keys (final mode)
    const float* kpos_f32;
    float inv_T, dq_scale;
    int round_bf16;
    float* loss_rows;            // [B]
    float* dq;                   // [B, D]
    int32_t* pos_is_max;         // [B]
    float* max_logit;            // [B] (nullable)
    float* loss_mean;            // [1] (nullable, with acc_pct)
    float* acc_pct;              // [1]
    float* packed;               // [B, D + 4] (packed mode)
    unsigned int* counters;      // [query tiles + 1], zero between launches
};
}  // namespace tc

// Programmatic dependent launch: a kernel launched through launch_pdl() may start while its predecessor in the stream is
// still draining; it must call pdl_wait() before touching anything a predecessor wrote (or reads -- WAR) and may call
// pdl_launch_dependents() as soon as its successor is allowed to start launching.  Both are no-ops for kernels launched
// without the attribute.  MOMA_B200_PDL=0 turns the attribute off (A/B switch, read once).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(static_cast<Args&&>(args))...);
    if (e != cudaSuccess) note_launch_error(e);
    return e;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// cp.async (global -> shared without a register stage); `valid == false` zero-fills the destination instead
__device__ __forceinline__ void cp_async16(float* dst, const float* src, bool valid) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src, bool valid) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit accesses (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_rw(const float4* p) {   // data we are about to overwrite
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- peer-memory exchange protocol (csrc/peer.cu; also spoken by the InfoNCE merge / combine kernels, csrc/nce_simt.cu)
constexpr int kPeerMaxWorld = 16, kPeerChannels = 4, kPeerThreads = 256;
// A peer may legitimately be late by a host-side hiccup (graph instantiation, allocator growth): wait ~30 s of SM clocks
// before declaring it dead -- long enough for that, short enough that a lost rank becomes an error instead of a hang.
constexpr long long kPeerTimeoutClk = 60000000000ll;

struct PeerCtrl {
    unsigned long long epoch[kPeerChannels];
    unsigned int ticket_done[kPeerChannels];
};
// One channel of the symmetric allocation as a kernel argument (bases: device array of every rank's base address)
struct PeerLink {
    const unsigned long long* bases;
    long long ctrl_off, data_off, region_bytes;
    int rank, world, channel, rows_per_rank;
};
#ifdef __CUDACC__
__device__ __forceinline__ uint4 ld_volatile16(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
#endif

}  // namespace moma
