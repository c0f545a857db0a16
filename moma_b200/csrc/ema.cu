// Multi-tensor momentum (EMA) update -- HBM-bound, 12 B per element.
// Replaces learning/contrast_trainer.py:207-211:
//     p2.data.mul_(m).add_(p1.detach().data, alpha=(1 - m))
// One launch walks a chunk table covering every parameter tensor; 128-bit
// streaming loads/stores; exactly the reference's two fp32 roundings.
#include "common.cuh"
#include <stdlib.h>

namespace moma {

constexpr int kEmaThreads = 256;
constexpr int kEmaVecPerThread = 8;                       // float4 per thread per array
constexpr int64_t kEmaChunk = (int64_t)kEmaThreads * kEmaVecPerThread * 4;   // 8192 elements

struct alignas(32) EmaChunk {
    const float* src;
    float* dst;
    int32_t count;       // elements in this chunk (<= kEmaChunk)
    int32_t vec_ok;      // both pointers 16-byte aligned
    int64_t pad;
};
static_assert(sizeof(EmaChunk) == 32, "EmaChunk layout");

__device__ __forceinline__ float ema1(float d, float s, float m, float a) {
    return __fmaf_rn(a, s, __fmul_rn(d, m));
}

// Persistent: a few CTAs per SM walk the chunk table with a grid stride.  Each thread keeps 16 independent 128-bit
// loads in flight (64 KB per CTA), which is enough to saturate HBM from ~1-2 CTAs per SM -- and it leaves the rest of
// every SM free, so the latency-bound kernels of the criterion step that run concurrently (other graph branches)
// get their CTAs scheduled at once instead of queueing behind a full-occupancy streaming grid.
__global__ void __launch_bounds__(kEmaThreads)
ema_multi_kernel(const EmaChunk* __restrict__ table, int n_chunks, float m, float a) {
    const int t = threadIdx.x;
  for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
    const EmaChunk c = table[ci];
    if (c.vec_ok) {
        const int nvec = c.count >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(c.src);
        float4* d4 = reinterpret_cast<float4*>(c.dst);
        float4 s[kEmaVecPerThread], d[kEmaVecPerThread];
#pragma unroll
        for (int i = 0; i < kEmaVecPerThread; ++i) {           // all loads first (MLP)
            const int v = t + i * kEmaThreads;
            if (v < nvec) { s[i] = ld_stream(s4 + v); d[i] = ld_rw(d4 + v); }
        }
#pragma unroll
        for (int i = 0; i < kEmaVecPerThread; ++i) {
            const int v = t + i * kEmaThreads;
            if (v < nvec) {
                float4 r;
                r.x = ema1(d[i].x, s[i].x, m, a); r.y = ema1(d[i].y, s[i].y, m, a);
                r.z = ema1(d[i].z, s[i].z, m, a); r.w = ema1(d[i].w, s[i].w, m, a);
                st_stream(d4 + v, r);
            }
        }
        const int tail = nvec << 2;
        if (t < c.count - tail) c.dst[tail + t] = ema1(c.dst[tail + t], c.src[tail + t], m, a);
    } else {
        for (int i = t; i < c.count; i += kEmaThreads) c.dst[i] = ema1(c.dst[i], c.src[i], m, a);
    }
  }
}

static int ema_ctas_per_sm() {
    static const int v = [] { const char* e = getenv("MOMA_B200_EMA_CTAS_PER_SM"); int n = e ? atoi(e) : 2; return n < 1 ? 1 : (n > 8 ? 8 : n); }();
    return v;
}

}  // namespace moma

using namespace moma;

extern "C" __attribute__((visibility("default"))) int moma_ema_plan_size(int n_tensors, const int64_t* numels, int64_t* n_chunks,
                                  size_t* table_bytes) {
    MOMA_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || numels) && n_chunks && table_bytes,
                 MOMA_ERR_INVALID, "ema_plan_size: bad arguments");
    int64_t chunks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        MOMA_REQUIRE(numels[i] >= 0, MOMA_ERR_INVALID, "ema_plan_size: negative numel");
        chunks += (numels[i] + kEmaChunk - 1) / kEmaChunk;
    }
    *n_chunks = chunks;
    *table_bytes = (size_t)chunks * sizeof(EmaChunk);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_ema_plan_fill(int n_tensors, const void* const* src_ptrs, void* const* dst_ptrs,
                                  const int64_t* numels, void* host_table, size_t table_bytes) {
    MOMA_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (src_ptrs && dst_ptrs && numels)),
                 MOMA_ERR_INVALID, "ema_plan_fill: bad arguments");
    int64_t need = 0; size_t bytes = 0;
    int rc = moma_ema_plan_size(n_tensors, numels, &need, &bytes);
    if (rc != MOMA_OK) return rc;
    MOMA_REQUIRE(table_bytes >= bytes && (bytes == 0 || host_table), MOMA_ERR_WORKSPACE,
                 "ema_plan_fill: table too small (%zu < %zu)", table_bytes, bytes);
    EmaChunk* t = static_cast<EmaChunk*>(host_table);
    int64_t c = 0;
    for (int i = 0; i < n_tensors; ++i) {
        if (numels[i] == 0) continue;
        MOMA_REQUIRE(src_ptrs[i] && dst_ptrs[i], MOMA_ERR_INVALID, "ema_plan_fill: null tensor %d", i);
        MOMA_REQUIRE((reinterpret_cast<uintptr_t>(src_ptrs[i]) & 3u) == 0 &&
                     (reinterpret_cast<uintptr_t>(dst_ptrs[i]) & 3u) == 0, MOMA_ERR_ALIGN,
                     "ema_plan_fill: tensor %d not 4-byte aligned", i);
        const float* s = static_cast<const float*>(src_ptrs[i]);
        float* d = static_cast<float*>(dst_ptrs[i]);
        const int vec_ok = aligned16(s) && aligned16(d);
        for (int64_t off = 0; off < numels[i]; off += kEmaChunk, ++c) {
            int64_t cnt = numels[i] - off; if (cnt > kEmaChunk) cnt = kEmaChunk;
            t[c].src = s + off; t[c].dst = d + off; t[c].count = (int32_t)cnt; t[c].vec_ok = vec_ok; t[c].pad = 0;
        }
    }
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_ema_multi(const void* dev_table, int64_t n_chunks, float m, float one_minus_m,
                              moma_stream_t stream) {
    MOMA_REQUIRE(n_chunks >= 0, MOMA_ERR_INVALID, "ema_multi: negative chunk count");
    if (n_chunks == 0) return MOMA_OK;
    MOMA_REQUIRE(dev_table && aligned16(dev_table), MOMA_ERR_ALIGN, "ema_multi: table null/unaligned");
    MOMA_REQUIRE(n_chunks < (1ll << 31), MOMA_ERR_UNSUPPORTED, "ema_multi: too many chunks");
    int64_t grid = (int64_t)sm_count() * ema_ctas_per_sm();
    if (grid > n_chunks) grid = n_chunks;
    ema_multi_kernel<<<(unsigned)grid, kEmaThreads, 0, as_stream(stream)>>>(
        static_cast<const EmaChunk*>(dev_table), (int)n_chunks, m, one_minus_m);
    MOMA_CUDA_LAUNCH_CHECK("ema_multi");
    note_launches(1);
    return MOMA_OK;
}
