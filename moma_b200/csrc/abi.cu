// C-ABI plumbing: version, error string, device probe.
#include <stdarg.h>
#include "common.cuh"
#include <stdlib.h>

#include <atomic>
namespace moma {

static std::atomic<long long> g_launches{0};
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("MOMA_B200_PDL"); return !(e != nullptr && e[0] == '0'); }();
    return on;
}
bool use_simt_gemm() {
    static const bool simt = [] { const char* e = getenv("MOMA_B200_GEMM"); return e != nullptr && std::string(e) == "simt"; }();
    return simt;
}
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace moma

extern "C" __attribute__((visibility("default"))) long long moma_debug_launch_count(int reset) {
    return reset ? moma::g_launches.exchange(0) : moma::g_launches.load();
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
