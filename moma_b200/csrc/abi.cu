// C-ABI plumbing: version, error string, device probe.
#include <stdarg.h>
#include "common.cuh"
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <set>
#include <utility>
namespace moma {

static thread_local cudaError_t g_launch_err = cudaSuccess;
void note_launch_error(cudaError_t e) { if (g_launch_err == cudaSuccess) g_launch_err = e; }
cudaError_t take_launch_error() { const cudaError_t e = g_launch_err; g_launch_err = cudaSuccess; return e; }

cudaError_t ensure_dyn_smem(const void* func, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if (done.count({func, dev})) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.insert({func, dev});
    else note_launch_error(e);
    return e;
}

static std::atomic<long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::mutex g_flops_mu;
static double g_flops[4] = {0, 0, 0, 0};
void note_flops(int kind, double flops) {
    if (kind < 0 || kind >= 4) return;
    std::lock_guard<std::mutex> lock(g_flops_mu);
    g_flops[kind] += flops;
}

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
static std::atomic<int> g_pdl{-1};     // -1: not decided yet (MOMA_B200_PDL read on first use), else 0 / 1
bool pdl_enabled() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("MOMA_B200_PDL");
        v = !(e != nullptr && e[0] == '0');
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}
bool use_simt_gemm() {
    static const bool simt = [] { const char* e = getenv("MOMA_B200_GEMM"); return e != nullptr && std::string(e) == "simt"; }();
    return simt;
}
int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

}  // namespace moma

extern "C" __attribute__((visibility("default"))) long long moma_debug_launch_count(int reset) {
    return reset ? moma::g_launches.exchange(0) : moma::g_launches.load();
}
extern "C" __attribute__((visibility("default"))) int moma_debug_set_pdl(int enable) {
    const bool was = moma::pdl_enabled();
    moma::g_pdl.store(enable ? 1 : 0, std::memory_order_relaxed);
    return was ? 1 : 0;
}
extern "C" __attribute__((visibility("default"))) double moma_debug_flops(int kind, int reset) {
    if (kind < 0 || kind >= 4) return 0.0;
    std::lock_guard<std::mutex> lock(moma::g_flops_mu);
    const double v = moma::g_flops[kind];
    if (reset) moma::g_flops[kind] = 0.0;
    return v;
}
extern "C" __attribute__((visibility("default"))) int moma_abi_version(void) { return MOMA_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) const char* moma_last_error(void) { return moma::g_err; }

extern "C" __attribute__((visibility("default"))) int moma_has_tcgen05(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError(); return 0;
    }
    return major == 10 ? 1 : 0;
}
