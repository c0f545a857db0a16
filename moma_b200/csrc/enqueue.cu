// L2-normalise, ring enqueue into the momentum memory bank, queue pointer.
// HBM-bound row kernels with 128-bit accesses.
// Replaces MoMA/criterion_moco_att.py:12-18 (Normalize) and
// MoMA/mem_moco.py:14-27 (_update_pointer / _update_memory).
#include "common.cuh"

namespace moma {

// ---------------------------------------------------------------- Normalize
// One warp per row; the row is read once into registers when D <= 32*4*kMaxVec.
constexpr int kNormWarps = 4;

template <bool kBackward>
__global__ void __launch_bounds__(kNormWarps * 32)
l2norm_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ out,
              int64_t rows, int64_t D, float eps) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kNormWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* x4 = reinterpret_cast<const float4*>(x + row * D);
    const float4* g4 = kBackward ? reinterpret_cast<const float4*>(g + row * D) : nullptr;
    float4* o4 = reinterpret_cast<float4*>(out + row * D);
    const int nvec = (int)(D >> 2);
    float ss = 0.f, dot = 0.f;
    for (int v = lane; v < nvec; v += 32) {
        const float4 a = x4[v];
        ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        if (kBackward) {
            const float4 b = g4[v];
            dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        }
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float den = fmaxf(nrm, eps);
    if (!kBackward) {
        for (int v = lane; v < nvec; v += 32) {
            float4 a = x4[v];
            a.x = a.x / den; a.y = a.y / den; a.z = a.z / den; a.w = a.w / den;
            o4[v] = a;
        }
    } else {
        // d/dx [x / clamp_min(||x||, eps)] . g  =  g/den - live * x * (x.g) / (den^2 * ||x||)
        dot = warp_sum(dot);
        const float live = (nrm >= eps) ? 1.f : 0.f;
        const float coef = (nrm > 0.f) ? live * dot / (den * den * nrm) : 0.f;
        for (int v = lane; v < nvec; v += 32) {
            const float4 a = x4[v];
            const float4 b = g4[v];
            float4 r;
            r.x = b.x / den - a.x * coef; r.y = b.y / den - a.y * coef;
            r.z = b.z / den - a.z * coef; r.w = b.w / den - a.w * coef;
            o4[v] = r;
        }
    }
}

// ------------------------------------------------------------------ enqueue
// One warp per key row j: global id g = (index + j) % K (int64, exactly the
// reference's fmod(arange(n) + index, K)); owner rank g % W, local slot g / W.
// Writes the fp32 master row verbatim and the bf16 shadow row.
constexpr int kEnqWarps = 4;

__device__ __forceinline__ uint2 pack_bf16x4(const float4& a) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a.x, a.y);
    __nv_bfloat162 hi = __floats2bfloat162_rn(a.z, a.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&lo);
    r.y = *reinterpret_cast<uint32_t*>(&hi);
    return r;
}

__global__ void __launch_bounds__(kEnqWarps * 32)
enqueue_kernel(const float* __restrict__ keys, int64_t n, int64_t D, float* __restrict__ queue,
               __nv_bfloat16* __restrict__ shadow, int64_t K, int64_t index,
               const int64_t* __restrict__ index_dev, int rank, int world, int normalize, float eps,
               int64_t key_start, int64_t key_stride) {
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * kEnqWarps + (threadIdx.x >> 5);
    if (j >= n) return;
    if (index_dev) index = *index_dev;
    // provided row j is row key_start + j * key_stride of the step's key list (identity by default)
    const int64_t gid = (key_start + j * key_stride + index) % K;
    if (gid % world != rank) return;
    const int64_t slot = gid / world;
    const float4* s4 = reinterpret_cast<const float4*>(keys + j * D);
    float4* q4 = reinterpret_cast<float4*>(queue + slot * D);
    uint2* b2 = shadow ? reinterpret_cast<uint2*>(shadow + slot * D) : nullptr;
    const int nvec = (int)(D >> 2);
    float den = 1.f;
    if (normalize) {
        float ss = 0.f;
        for (int v = lane; v < nvec; v += 32) {
            const float4 a = s4[v];
            ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
        den = fmaxf(sqrtf(warp_sum(ss)), eps);
    }
    for (int v = lane; v < nvec; v += 32) {
        float4 a = s4[v];
        if (normalize) { a.x = a.x / den; a.y = a.y / den; a.z = a.z / den; a.w = a.w / den; }
        q4[v] = a;
        if (b2) b2[v] = pack_bf16x4(a);
    }
}

__global__ void enqueue_ids_kernel(int64_t n, int64_t index, const int64_t* __restrict__ index_dev,
                                   int64_t K, int64_t* __restrict__ out) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (index_dev) index = *index_dev;
    if (j < n) out[j] = (j + index) % K;
}

__global__ void pointer_advance_kernel(int64_t* index_dev, int64_t n, int64_t K) {
    pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x == 0 && blockIdx.x == 0) *index_dev = (*index_dev + n) % K;
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t nvec,
                 int64_t numel) {
    pdl_wait();
    pdl_launch_dependents();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    uint2* d2 = reinterpret_cast<uint2*>(dst);
    for (; v < nvec; v += stride) d2[v] = pack_bf16x4(ld_stream(s4 + v));
    // scalar tail
    const int64_t tail = nvec << 2;
    const int64_t t = tail + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < numel) dst[t] = __float2bfloat16_rn(src[t]);
}

}  // namespace moma

using namespace moma;

static int check_rows(const char* who, const void* a, const void* b, int64_t rows, int64_t D) {
    MOMA_REQUIRE(rows >= 0 && D > 0, MOMA_ERR_INVALID, "%s: bad shape [%lld, %lld]", who,
                 (long long)rows, (long long)D);
    MOMA_REQUIRE(D % 4 == 0, MOMA_ERR_UNSUPPORTED, "%s: D=%lld must be a multiple of 4", who, (long long)D);
    MOMA_REQUIRE(rows == 0 || (a && b), MOMA_ERR_INVALID, "%s: null pointer", who);
    MOMA_REQUIRE(aligned16(a) && aligned16(b), MOMA_ERR_ALIGN, "%s: pointers must be 16-byte aligned", who);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_l2norm_fwd(const float* x, float* y, int64_t rows, int64_t D, float eps,
                               moma_stream_t stream) {
    int rc = check_rows("l2norm_fwd", x, y, rows, D);
    if (rc != MOMA_OK || rows == 0) return rc;
    const unsigned grid = (unsigned)((rows + kNormWarps - 1) / kNormWarps);
    launch_pdl(l2norm_kernel<false>, dim3(grid), dim3(kNormWarps * 32), 0, as_stream(stream), x, nullptr, y, rows, D, eps);
    MOMA_CUDA_LAUNCH_CHECK("l2norm_fwd");
    note_launches(1);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_l2norm_bwd(const float* x, const float* grad_y, float* grad_x, int64_t rows,
                               int64_t D, float eps, moma_stream_t stream) {
    int rc = check_rows("l2norm_bwd", x, grad_x, rows, D);
    if (rc != MOMA_OK || rows == 0) return rc;
    MOMA_REQUIRE(grad_y && aligned16(grad_y), MOMA_ERR_ALIGN, "l2norm_bwd: grad_y null/unaligned");
    const unsigned grid = (unsigned)((rows + kNormWarps - 1) / kNormWarps);
    launch_pdl(l2norm_kernel<true>, dim3(grid), dim3(kNormWarps * 32), 0, as_stream(stream), x, grad_y, grad_x, rows, D, eps);
    MOMA_CUDA_LAUNCH_CHECK("l2norm_bwd");
    note_launches(1);
    return MOMA_OK;
}

static int enqueue_impl(const float* keys, int64_t n, int64_t D, float* queue_f32,
                        void* queue_bf16, int64_t K, int64_t index, const int64_t* index_dev,
                        int shard_rank, int shard_world, int normalize, float eps, int64_t key_start,
                        int64_t key_stride, moma_stream_t stream) {
    int rc = check_rows("enqueue", keys, queue_f32, n, D);
    if (rc != MOMA_OK) return rc;
    MOMA_REQUIRE(K > 0 && shard_world >= 1 && shard_rank >= 0 && shard_rank < shard_world,
                 MOMA_ERR_INVALID, "enqueue: bad K/shard (%lld, %d/%d)", (long long)K, shard_rank, shard_world);
    MOMA_REQUIRE(K % shard_world == 0, MOMA_ERR_INVALID, "enqueue: K=%lld not divisible by shard_world=%d",
                 (long long)K, shard_world);
    MOMA_REQUIRE(key_start >= 0 && key_stride >= 1 && key_start + (n > 0 ? n - 1 : 0) * key_stride < K, MOMA_ERR_INVALID,
                 "enqueue: n=%lld > K=%lld gives duplicate ids (undefined in the reference)",
                 (long long)(key_start + n * key_stride), (long long)K);
    MOMA_REQUIRE(index_dev || (index >= 0 && index < K), MOMA_ERR_INVALID, "enqueue: index out of range");
    MOMA_REQUIRE(!queue_bf16 || (D % 8 == 0 && aligned16(queue_bf16)), MOMA_ERR_ALIGN,
                 "enqueue: bf16 shadow needs D %% 8 == 0 and 16-byte alignment");
    if (n == 0) return MOMA_OK;
    const unsigned grid = (unsigned)((n + kEnqWarps - 1) / kEnqWarps);
    launch_pdl(enqueue_kernel, dim3(grid), dim3(kEnqWarps * 32), 0, as_stream(stream),
               keys, n, D, queue_f32, static_cast<__nv_bfloat16*>(queue_bf16), K, index, index_dev,
               shard_rank, shard_world, normalize, eps, key_start, key_stride);
    MOMA_CUDA_LAUNCH_CHECK("enqueue");
    note_launches(1);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_enqueue(const float* keys, int64_t n, int64_t D, float* queue_f32,
                            void* queue_bf16, int64_t K, int64_t index, const int64_t* index_dev,
                            int shard_rank, int shard_world, int normalize, float eps,
                            moma_stream_t stream) {
    return enqueue_impl(keys, n, D, queue_f32, queue_bf16, K, index, index_dev, shard_rank, shard_world, normalize,
                        eps, 0, 1, stream);
}

// keys[i] is row key_start + i * key_stride of the step's key list (a rank that only computed the rows it owns)
extern "C" __attribute__((visibility("default"))) int moma_enqueue_strided(const float* keys, int64_t n, int64_t D,
                            float* queue_f32, void* queue_bf16, int64_t K, int64_t index, const int64_t* index_dev,
                            int shard_rank, int shard_world, int64_t key_start, int64_t key_stride,
                            moma_stream_t stream) {
    return enqueue_impl(keys, n, D, queue_f32, queue_bf16, K, index, index_dev, shard_rank, shard_world, 0, 1e-12f,
                        key_start, key_stride, stream);
}

extern "C" __attribute__((visibility("default"))) int moma_enqueue_ids(int64_t n, int64_t index, const int64_t* index_dev, int64_t K,
                                int64_t* out_ids, moma_stream_t stream) {
    MOMA_REQUIRE(n >= 0 && K > 0 && (n == 0 || out_ids), MOMA_ERR_INVALID, "enqueue_ids: bad arguments");
    if (n == 0) return MOMA_OK;
    launch_pdl(enqueue_ids_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, as_stream(stream), n, index, index_dev, K, out_ids);
    MOMA_CUDA_LAUNCH_CHECK("enqueue_ids");
    note_launches(1);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_pointer_advance(int64_t* index_dev, int64_t n, int64_t K, moma_stream_t stream) {
    MOMA_REQUIRE(index_dev && n >= 0 && K > 0, MOMA_ERR_INVALID, "pointer_advance: bad arguments");
    launch_pdl(pointer_advance_kernel, dim3(1), dim3(32), 0, as_stream(stream), index_dev, n, K);
    MOMA_CUDA_LAUNCH_CHECK("pointer_advance");
    note_launches(1);
    return MOMA_OK;
}

extern "C" __attribute__((visibility("default"))) int moma_cast_bf16(const float* src, void* dst_bf16, int64_t numel, moma_stream_t stream) {
    MOMA_REQUIRE(numel >= 0 && (numel == 0 || (src && dst_bf16)), MOMA_ERR_INVALID, "cast_bf16: bad arguments");
    MOMA_REQUIRE(aligned16(src) && aligned16(dst_bf16), MOMA_ERR_ALIGN, "cast_bf16: unaligned pointers");
    if (numel == 0) return MOMA_OK;
    const int64_t nvec = numel >> 2;
    int64_t blocks = (nvec + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    launch_pdl(cast_bf16_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream),
               src, static_cast<__nv_bfloat16*>(dst_bf16), nvec, numel);
    MOMA_CUDA_LAUNCH_CHECK("cast_bf16");
    note_launches(1);
    return MOMA_OK;
}

// y = x * (*scalar): the chain-rule multiply of the fused InfoNCE gradient by the upstream scalar (d total / d loss_kd,
// a device-resident 0-dim tensor -- opt.beta times whatever follows, helper/loops_moma.py:345) without a host sync and
// inside the step's chain of programmatic dependent launches.
__global__ void __launch_bounds__(256) scale_by_scalar_kernel(const float* __restrict__ x, const float* __restrict__ scalar,
                                                              float* __restrict__ y, int64_t nvec, int64_t numel) {
    pdl_wait();
    pdl_launch_dependents();
    const float s = *scalar;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float4 v = reinterpret_cast<const float4*>(x)[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        reinterpret_cast<float4*>(y)[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (numel & 3)) y[(nvec << 2) + threadIdx.x] = x[(nvec << 2) + threadIdx.x] * s;
}
extern "C" __attribute__((visibility("default"))) int moma_scale_by_scalar(const float* x, const float* scalar_dev, float* y, int64_t numel,
                                                                          moma_stream_t stream) {
    MOMA_REQUIRE(numel >= 0 && (numel == 0 || (x && scalar_dev && y)), MOMA_ERR_INVALID, "scale_by_scalar: bad arguments");
    MOMA_REQUIRE(aligned16(x) && aligned16(y), MOMA_ERR_ALIGN, "scale_by_scalar: unaligned pointers");
    if (numel == 0) return MOMA_OK;
    const int64_t nvec = numel >> 2;
    int64_t blocks = (nvec + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    launch_pdl(scale_by_scalar_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), x, scalar_dev, y, nvec, numel);
    MOMA_CUDA_LAUNCH_CHECK("scale_by_scalar");
    note_launches(1);
    return MOMA_OK;
}
