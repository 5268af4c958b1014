// Library-wide plumbing: thread-local error message, device queries.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace cmh {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace cmh

namespace cmh { unsigned long long launches(); }
extern "C" int cmh_abi_version(void) { return CMH_ABI_VERSION; }
extern "C" unsigned long long cmh_launch_count(void) { return cmh::launches(); }
extern "C" const char* cmh_last_error(void) { return cmh::g_err; }

extern "C" int cmh_device_info(int* sm, int* cc_major, int* cc_minor, uint64_t* total_mem) {
    int dev = 0;
    CMH_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    CMH_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm) *sm = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = (uint64_t)p.totalGlobalMem;
    CMH_REQUIRE(p.major == 10, CMH_ERR_DEVICE, "cmh_b200 is built for sm_100a only; device %d is sm_%d%d", dev, p.major,
                p.minor);
    return CMH_OK;
}
