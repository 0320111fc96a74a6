// Version, error reporting and device check of the C ABI.
#include "abt_internal.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace abt {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_device_sm100() {
    static thread_local int cached_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    if (dev == cached_dev) return 0;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return set_error(ABT_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    if (major != 10) return set_error(ABT_ERR_DEVICE, "device %d has compute capability %d.x; libabt_b200 is sm_100a only (no fallback)", dev, major);
    cached_dev = dev;
    return 0;
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace abt

extern "C" long long abt_debug_launch_count(int reset) {
    return reset ? abt::g_launches.exchange(0) : abt::g_launches.load();
}

extern "C" int abt_version(void) { return ABT_VERSION; }
extern "C" const char* abt_last_error(void) { return abt::g_err; }
extern "C" int abt_device_check(void) { return abt::check_device_sm100(); }
