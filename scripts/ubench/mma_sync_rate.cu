// Micro-benchmark: warp-level mma.sync issue rates on sm_100a (legacy tensor-core path), 1 and 2 warps per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float d[4][4] = {};
    uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
    float fa[8];
    for (int j = 0; j < 8; ++j) fa[j] = threadIdx.x * 0.001f + j;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {      // 4 independent accumulator chains
            if (MODE == 0) asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                        : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            if (MODE == 1) asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                        : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            if (MODE == 2) asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                                        : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
            if (MODE == 3) {               // 64 FFMA: the SIMT equivalent of a quarter of an m16n8k8 (1024 MAC / 32 lanes = 32 FMA per lane)
#pragma unroll
                for (int e = 0; e < 4; ++e) { d[j][e] = fmaf(fa[e], fa[e + 4], d[j][e]); d[j][e] = fmaf(fa[e + 1], fa[(e + 5) & 7], d[j][e]); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < 4; ++j) for (int e = 0; e < 4; ++e) s += d[j][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    const int iters = 4000;
    const char* names[] = {"mma.sync m16n8k8  tf32 (1024 MAC)", "mma.sync m16n8k16 bf16 (2048 MAC)", "mma.sync m16n8k4  tf32 (512 MAC)", "8 FFMA per chain x4 (32 FMA/lane)"};
    for (int warps = 4; warps <= 8; warps += 4)
        for (int mode = 0; mode < 4; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
                if (mode == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
                if (mode == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
                if (mode == 3) k<3><<<148, warps * 32>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%d warps/SM  %-36s %7.1f clk per 4 instructions per warp\n", warps, names[mode], (double)c / iters);
        }
    return 0;
}
