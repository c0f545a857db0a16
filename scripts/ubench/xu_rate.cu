// Micro-benchmark: issue rate of MUFU.EX2 and F2FP (fp32x2 -> bf16x2 pack) on sm_100a.
// 8 warps per SM (2 per scheduler), 16 independent dependency chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, int iters) {
    float x[16];
    for (int j = 0; j < 16; ++j) x[j] = -(threadIdx.x * 1e-3f + j * 0.01f);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            if (MODE == 0) { x[j] = ex2(x[j]); x[j + 1] = ex2(x[j + 1]); }
            if (MODE == 1) { uint32_t r = pack(x[j], x[j + 1]); x[j] = __uint_as_float(r); }
            if (MODE == 2) { float a = ex2(x[j]), b = ex2(x[j + 1]); uint32_t r = pack(a, b); x[j] = __uint_as_float(r); x[j + 1] = a; }
            if (MODE == 3) { uint32_t a = (__float_as_uint(x[j]) + 0x8000u), b = (__float_as_uint(x[j + 1]) + 0x8000u);
                             x[j] = __uint_as_float(__byte_perm(a, b, 0x7632)); }
            if (MODE == 4) { float a = ex2(x[j]), b = ex2(x[j + 1]);
                             uint32_t ua = (__float_as_uint(a) + 0x8000u), ub = (__float_as_uint(b) + 0x8000u);
                             x[j] = __uint_as_float(__byte_perm(ua, ub, 0x7632)); x[j + 1] = a; }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < 16; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    const char* names[] = {"16 MUFU.EX2", "8 F2FP.BF16.PACK", "16 EX2 + 8 F2FP", "8 x (2 IADD + PRMT)", "16 EX2 + 8 x (2 IADD + PRMT)"};
    for (int mode = 0; mode < 5; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            if (mode == 0) k<0><<<148, 256>>>(out, cyc, iters);
            if (mode == 1) k<1><<<148, 256>>>(out, cyc, iters);
            if (mode == 2) k<2><<<148, 256>>>(out, cyc, iters);
            if (mode == 3) k<3><<<148, 256>>>(out, cyc, iters);
            if (mode == 4) k<4><<<148, 256>>>(out, cyc, iters);
            cudaDeviceSynchronize();
        }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-32s %8.1f clk per loop iteration, 2 warps per scheduler\n", names[mode], (double)c / iters);
    }
    return 0;
}
