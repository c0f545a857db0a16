// Micro-benchmark: issue rates of the softmax instruction mix on sm_100a (8 warps / SM = 2 per scheduler).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)), "l"(*reinterpret_cast<uint64_t*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float fadd(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float d; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, int iters, float s) {
    float2 x[16];
    for (int j = 0; j < 16; ++j) x[j] = make_float2(-(threadIdx.x * 1e-3f + j * 0.01f), -(threadIdx.x * 2e-3f + j * 0.02f));
    const float2 sc = make_float2(s, s), ng = make_float2(-s * 0.01f, -s * 0.01f);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0) x[j] = ffma2(x[j], sc, ng);
            if (MODE == 1) x[j] = fadd2(x[j], ng);
            if (MODE == 2) { x[j].x = ffma(x[j].x, s, ng.x); x[j].y = ffma(x[j].y, s, ng.y); }
            if (MODE == 3) { x[j].x = fadd(x[j].x, ng.x); x[j].y = fadd(x[j].y, ng.y); }
            if (MODE == 4) x[j].x = fmax3(x[j].x, x[j].y, ng.x);
            if (MODE == 5) { float2 t = ffma2(x[j], sc, ng); x[j] = fadd2(x[j], t); }
            if (MODE == 6) {   // the kernel's per-pair work: scale, 2 exps, sum, pack
                float2 t = ffma2(x[j], sc, ng); t.x = ex2(t.x); t.y = ex2(t.y); x[j] = fadd2(x[j], t);
                x[(j + 1) & 15].x = __uint_as_float(pack(t.x, t.y));
            }
            if (MODE == 7) {   // same with scalar fp32 instead of the packed forms
                float a = ffma(x[j].x, s, ng.x), b = ffma(x[j].y, s, ng.y); a = ex2(a); b = ex2(b);
                x[j].x = fadd(x[j].x, a); x[j].y = fadd(x[j].y, b); x[(j + 1) & 15].x = __uint_as_float(pack(a, b));
            }
        }
    }
    long long t1 = clock64();
    float r = 0.f;
    for (int j = 0; j < 16; ++j) r += x[j].x + x[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    const char* names[] = {"16 FFMA2", "16 FADD2", "32 FFMA", "32 FADD", "16 FMNMX3", "16 FFMA2 + 16 FADD2",
                           "16 x (FFMA2, 2 MUFU, FADD2, F2FP)", "16 x (2 FFMA, 2 MUFU, 2 FADD, F2FP)"};
    for (int mode = 0; mode < 8; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (mode) {
                case 0: k<0><<<148, 256>>>(out, cyc, iters, 0.999f); break; case 1: k<1><<<148, 256>>>(out, cyc, iters, 0.999f); break;
                case 2: k<2><<<148, 256>>>(out, cyc, iters, 0.999f); break; case 3: k<3><<<148, 256>>>(out, cyc, iters, 0.999f); break;
                case 4: k<4><<<148, 256>>>(out, cyc, iters, 0.999f); break; case 5: k<5><<<148, 256>>>(out, cyc, iters, 0.999f); break;
                case 6: k<6><<<148, 256>>>(out, cyc, iters, 0.999f); break; case 7: k<7><<<148, 256>>>(out, cyc, iters, 0.999f); break;
            }
            cudaDeviceSynchronize();
        }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-40s %8.1f clk per loop iteration (2 warps per scheduler)\n", names[mode], (double)c / iters);
    }
    return 0;
}
