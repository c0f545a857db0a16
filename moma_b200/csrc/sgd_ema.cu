// SGD(momentum, weight decay) + momentum-encoder EMA in ONE multi-tensor pass -- HBM-bound, 28 B per element
// (read p, g, buf, ema; write p, buf, ema) instead of 20 B (optimizer.step) + 12 B (momentum_update) in two passes with
// ~3 launches per tensor.  SURVEY 8f-2: train_student_moma.py:389-392 (optim.SGD), helper/loops_moma.py:359-361
// (optimizer.step) followed by the next step's trainer.momentum_update (loops_moma.py:309,
// learning/contrast_trainer.py:207-211), which reads the weights the optimizer has just written.
//
// Arithmetic = torch.optim.SGD (dampening 0, no Nesterov) then momentum_update, rounding for rounding:
//     d   = fma(wd, p, g)                       # grad.add(param, alpha=weight_decay)
//     buf = first ? d : fl(buf * mu) + d        # buf.mul_(momentum).add_(d)
//     p   = fma(-lr, buf, p)                    # param.add_(buf, alpha=-lr)
//     ema = fma(1 - m, p, fl(ema * m))          # p2.data.mul_(m).add_(p1.data, alpha=1 - m)
#include "common.cuh"

namespace moma {

constexpr int kSgdThreads = 256;
constexpr int kSgdVec = 4;                                        // float4 per thread per array (4 arrays in flight)
constexpr int64_t kSgdChunk = (int64_t)kSgdThreads * kSgdVec * 4; // 4096 elements

struct alignas(64) SgdChunk {
    float* p;
    const float* g;
    float* buf;
    float* ema;
    int32_t count;
    int32_t vec_ok;
    int64_t pad[3];
};
static_assert(sizeof(SgdChunk) == 64, "SgdChunk layout");

struct SgdHyper { float lr, mu, wd, m, a; int first; };

__device__ __forceinline__ void sgd_ema1(float& p, float g, float& buf, float& ema, const SgdHyper& h) {
    const float d = __fmaf_rn(h.wd, p, g);
    buf = h.first ? d : __fadd_rn(__fmul_rn(buf, h.mu), d);
    p = __fmaf_rn(-h.lr, buf, p);
    ema = __fmaf_rn(h.a, p, __fmul_rn(ema, h.m));
}

__global__ void __launch_bounds__(kSgdThreads)
sgd_ema_multi_kernel(const SgdChunk* __restrict__ table, int n_chunks, const SgdHyper h) {
    const int t = threadIdx.x;
    for (int ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
        const SgdChunk c = table[ci];
        if (c.vec_ok) {
            const int nvec = c.count >> 2;
            float4* p4 = reinterpret_cast<float4*>(c.p);
            const float4* g4 = reinterpret_cast<const float4*>(c.g);
            float4* b4 = reinterpret_cast<float4*>(c.buf);
            float4* e4 = reinterpret_cast<float4*>(c.ema);
            float4 p[kSgdVec], g[kSgdVec], b[kSgdVec], e[kSgdVec];
#pragma unroll
            for (int i = 0; i < kSgdVec; ++i) {                   // all loads first
                const int v = t + i * kSgdThreads;
                if (v < nvec) {
                    p[i] = ld_rw(p4 + v); g[i] = ld_stream(g4 + v); e[i] = ld_rw(e4 + v);
                    b[i] = h.first ? make_float4(0.f, 0.f, 0.f, 0.f) : ld_rw(b4 + v);
                }
            }
#pragma unroll
            for (int i = 0; i < kSgdVec; ++i) {
                const int v = t + i * kSgdThreads;
                if (v < nvec) {
                    sgd_ema1(p[i].x, g[i].x, b[i].x, e[i].x, h); sgd_ema1(p[i].y, g[i].y, b[i].y, e[i].y, h);
                    sgd_ema1(p[i].z, g[i].z, b[i].z, e[i].z, h); sgd_ema1(p[i].w, g[i].w, b[i].w, e[i].w, h);
                    st_stream(p4 + v, p[i]); st_stream(b4 + v, b[i]); st_stream(e4 + v, e[i]);
                }
            }
            const int tail = nvec << 2;
            if (t < c.count - tail) {
                float pp = c.p[tail + t], bb = h.first ? 0.f : c.buf[tail + t], ee = c.ema[tail + t];
                sgd_ema1(pp, c.g[tail + t], bb, ee, h);
                c.p[tail + t] = pp; c.buf[tail + t] = bb; c.ema[tail + t] = ee;
            }
        } else {
            for (int i = t; i < c.count; i += kSgdThreads) {
                float pp = c.p[i], bb = h.first ? 0.f : c.buf[i], ee = c.ema[i];
                sgd_ema1(pp, c.g[i], bb, ee, h);
                c.p[i] = pp; c.buf[i] = bb; c.ema[i] = ee;
            }
        }
    }
}

}  // namespace moma

using namespace moma;

extern "C" __attribute__((visibility("default"))) int moma_sgd_ema_plan_size(int n_tensors, const int64_t* numels, int64_t* n_chunks,
                                                                            size_t* table_bytes) {
    MOMA_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || numels) && n_chunks && table_bytes, MOMA_ERR_INVALID,
                 "sgd_ema_plan_size: bad arguments");
    int64_t chunks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        MOMA_REQUIRE(numels[i] >= 0, MOMA_ERR_INVALID, "sgd_ema_plan_size: negative numel");
        chunks += (numels[i] + kSgdChunk - 1) / kSgdChunk;
    }
    *n_chunks = chunks;
    *table_bytes = (size_t)chunks * sizeof(SgdChunk);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_sgd_ema_plan_fill(int n_tensors, void* const* p_ptrs, const void* const* g_ptrs,
                                                                            void* const* buf_ptrs, void* const* ema_ptrs,
                                                                            const int64_t* numels, void* host_table, size_t table_bytes) {
    MOMA_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (p_ptrs && g_ptrs && buf_ptrs && ema_ptrs && numels)), MOMA_ERR_INVALID,
                 "sgd_ema_plan_fill: bad arguments");
    int64_t need = 0; size_t bytes = 0;
    int rc = moma_sgd_ema_plan_size(n_tensors, numels, &need, &bytes);
    if (rc != MOMA_OK) return rc;
    MOMA_REQUIRE(table_bytes >= bytes && (bytes == 0 || host_table), MOMA_ERR_WORKSPACE, "sgd_ema_plan_fill: table too small");
    SgdChunk* t = static_cast<SgdChunk*>(host_table);
    int64_t c = 0;
    for (int i = 0; i < n_tensors; ++i) {
        if (numels[i] == 0) continue;
        const void* ptrs[4] = {p_ptrs[i], g_ptrs[i], buf_ptrs[i], ema_ptrs[i]};
        bool vec = true;
        for (const void* q : ptrs) {
            MOMA_REQUIRE(q != nullptr, MOMA_ERR_INVALID, "sgd_ema_plan_fill: null tensor %d", i);
            MOMA_REQUIRE((reinterpret_cast<uintptr_t>(q) & 3u) == 0, MOMA_ERR_ALIGN, "sgd_ema_plan_fill: tensor %d not 4-byte aligned", i);
            vec = vec && aligned16(q);
        }
        for (int64_t off = 0; off < numels[i]; off += kSgdChunk, ++c) {
            int64_t cnt = numels[i] - off; if (cnt > kSgdChunk) cnt = kSgdChunk;
            t[c].p = static_cast<float*>(p_ptrs[i]) + off; t[c].g = static_cast<const float*>(g_ptrs[i]) + off;
            t[c].buf = static_cast<float*>(buf_ptrs[i]) + off; t[c].ema = static_cast<float*>(ema_ptrs[i]) + off;
            t[c].count = (int32_t)cnt; t[c].vec_ok = vec ? 1 : 0; t[c].pad[0] = t[c].pad[1] = t[c].pad[2] = 0;
        }
    }
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_sgd_ema_multi(const void* dev_table, int64_t n_chunks, float lr, float momentum,
                                                                        float weight_decay, int first_step, float m, float one_minus_m,
                                                                        moma_stream_t stream) {
    MOMA_REQUIRE(n_chunks >= 0, MOMA_ERR_INVALID, "sgd_ema_multi: negative chunk count");
    if (n_chunks == 0) return MOMA_OK;
    MOMA_REQUIRE(dev_table && aligned16(dev_table), MOMA_ERR_ALIGN, "sgd_ema_multi: table null/unaligned");
    MOMA_REQUIRE(n_chunks < (1ll << 31), MOMA_ERR_UNSUPPORTED, "sgd_ema_multi: too many chunks");
    int64_t grid = (int64_t)sm_count() * 2;
    if (grid > n_chunks) grid = n_chunks;
    SgdHyper h{lr, momentum, weight_decay, m, one_minus_m, first_step ? 1 : 0};
    sgd_ema_multi_kernel<<<(unsigned)grid, kSgdThreads, 0, as_stream(stream)>>>(static_cast<const SgdChunk*>(dev_table), (int)n_chunks, h);
    MOMA_CUDA_LAUNCH_CHECK("sgd_ema_multi");
    note_launches(1);
    return MOMA_OK;
}
