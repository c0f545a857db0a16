// Micro-benchmark: issue-to-completion rate of tcgen05.mma (kind::f16, bf16 -> f32, cta_group::1, M = 128) on sm_100a
// for the operand sources / layouts the InfoNCE kernel uses.  One CTA per SM, one issuing thread, no loads (smem zeroed).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// MODE 0: SS N=128 (A, B K-major)      1: SS N=256            2: TS N=128, B MN-major (the PV product)
//      3: TS N=128, B K-major          4: SS N=64             5: 8 x mode 0 then 8 x mode 2 alternating (the kernel's mix)
//      6: SS N=128, B MN-major         7: mode 5 but PV as N=256 in one accumulator (D=256 shape)
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int groups, int random_data) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar, bar2[3];
    __shared__ uint32_t tmem_base;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint32_t rng = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    auto rnd_bf16x2 = [&]() {       // two bf16 values in +-[0.5, 2)
        rng = rng * 1664525u + 1013904223u;
        return random_data ? ((rng & 0x80ff80ffu) | 0x3f003f00u) : 0u;
    };
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = rnd_bf16x2();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int j = 0; j < 3; ++j) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[j])));   // phases flip freely
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    for (int c = 0; c < 384; ++c) {     // the TMEM A operand (bf16 pairs)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((threadIdx.x >> 5) << 21) + c), "r"(rnd_bf16x2()) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        const uint64_t a_k = make_desc(base, 16, 1024);                 // A: 128 rows x 128 cols (2 x 64-col blocks of 16 KB)
        const uint64_t b_k = make_desc(base + 32768, 16, 1024);         // B K-major
        const uint64_t b_mn = make_desc(base + 32768, 128 * 128, 1024); // B MN-major (N contiguous), block stride 16 KB
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
#pragma unroll 1
            for (int ks = 0; ks < 8; ++ks) {
                const uint32_t koff = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32) >> 4;
                const uint32_t acc = ks > 0;
                if (MODE == 0) umma_ss(tmem, a_k + koff, b_k + koff, make_idesc(128, 128, 0, 0), acc);
                if (MODE == 1) umma_ss(tmem, a_k + koff, b_k + koff, make_idesc(128, 256, 0, 0), acc);
                if (MODE == 2) umma_ts(tmem + 384, tmem + ks * 8, b_mn + (uint64_t)(ks * (2048 >> 4)), make_idesc(128, 128, 0, 1), acc);
                if (MODE == 3) umma_ts(tmem + 384, tmem + ks * 8, b_k + koff, make_idesc(128, 128, 0, 0), acc);
                if (MODE == 4) umma_ss(tmem, a_k + koff, b_k + koff, make_idesc(128, 64, 0, 0), acc);
                if (MODE == 5 || MODE == 7 || MODE >= 8) umma_ss(tmem + (g % 3) * 128, a_k + koff, b_k + koff, make_idesc(128, 128, 0, 0), acc);
                if (MODE == 6) umma_ss(tmem, a_k + koff, b_mn + (uint64_t)(ks * (2048 >> 4)), make_idesc(128, 128, 0, 1), acc);
            }
            if (MODE >= 8) commit(smem_u32(&bar2[0]));
            if (MODE == 5 || MODE >= 8) {
#pragma unroll 1
                for (int ks = 0; ks < 8; ++ks)
                    umma_ts(tmem + 384, tmem + ((g + 1) % 3) * 128 + ks * 8, b_mn + (uint64_t)(ks * (2048 >> 4)),
                            make_idesc(128, 128, 0, 1), 1u);
            }
            if (MODE >= 8) commit(smem_u32(&bar2[1]));
            if (MODE >= 9) commit(smem_u32(&bar2[2]));
            if (MODE == 7) {
#pragma unroll 1
                for (int ks = 0; ks < 4; ++ks)
                    umma_ts(tmem + 256, tmem + ((g + 1) % 2) * 128 + ks * 8, b_mn + (uint64_t)(ks * (2048 >> 4)),
                            make_idesc(128, 256, 0, 1), 1u);
            }
        }
        commit(smem_u32(&bar));
        while (!try_wait(smem_u32(&bar), 0)) {}
        const long long t1 = clock64();
        if (blockIdx.x == 0) *cyc = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
template <int MODE>
static void run(const char* name, int mmas_per_group, double macs_per_group, long long* cyc, int random_data) {
    const int groups = 2000, smem = 161 * 1024;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float ms = 0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<148, 128, smem>>>(cyc, groups, random_data);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f clk/MMA  %6.0f MAC/clk/SM  %7.0f TFLOP/s chip (%.2f GHz)\n", name, (double)c / groups / mmas_per_group,
           macs_per_group * groups / c, 2.0 * macs_per_group * groups * 148 / (ms * 1e-3) / 1e12, c / (ms * 1e-3) / 1e9);
}
int main() {
    long long* cyc; cudaMalloc(&cyc, 8);
    const double m = 128.0 * 128 * 16;
    for (int rd = 0; rd < 2; ++rd) {
    printf("---- operands: %s\n", rd ? "random bf16" : "zeros");
    run<0>("SS  N=128  A,B K-major", 8, 8 * m, cyc, rd);
    run<1>("SS  N=256  A,B K-major", 8, 16 * m, cyc, rd);
    run<4>("SS  N=64   A,B K-major", 8, 4 * m, cyc, rd);
    run<6>("SS  N=128  B MN-major", 8, 8 * m, cyc, rd);
    run<2>("TS  N=128  B MN-major (PV)", 8, 8 * m, cyc, rd);
    run<3>("TS  N=128  B K-major", 8, 8 * m, cyc, rd);
    run<5>("8 SS N=128 + 8 TS N=128 alternating", 16, 16 * m, cyc, rd);
    run<7>("8 SS N=128 + 4 TS N=256 alternating", 12, 16 * m, cyc, rd);
    run<8>("8 SS + commit + 8 TS + commit", 16, 16 * m, cyc, rd);
    run<9>("8 SS + commit + 8 TS + 2 commits", 16, 16 * m, cyc, rd);
    }
    return 0;
}
