// Library-wide host plumbing: thread-local error string, device checks, version.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace mmb {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};

int check_launch(const char* what, int kernels) {
    g_launches.fetch_add(kernels, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return MMB_ECUDA;
    }
    return MMB_OK;
}

int pdl_mode() {
    static const int mode = [] {
        const char* e = getenv("MMB_PDL");
        const int m = e != nullptr ? atoi(e) : 0;      // default off: measured no gain on this path
        return m < 0 ? 0 : (m > 2 ? 2 : m);
    }();
    return mode;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// SMs left free by the persistent kernels (one CTA / CTA pair per SM: GEMM, attention) so that a concurrent NCCL
// kernel finds idle SMs instead of waiting for — and then delaying — a whole persistent wave.
static std::atomic<int> g_reserved{-1};
int persistent_sms() {
    int r = g_reserved.load(std::memory_order_relaxed);
    if (r < 0) {
        const char* e = getenv("MMB_RESERVE_SMS");
        r = e ? atoi(e) : 0;
        if (r < 0) r = 0;
        g_reserved.store(r, std::memory_order_relaxed);
    }
    int n = num_sms() - r;
    n &= ~1;                      // CTA pairs
    return n < 2 ? 2 : n;
}

}  // namespace mmb

extern "C" int mmb_set_reserved_sms(int n) {
    if (n < 0 || n > 64) {
        mmb::set_last_error("mmb_set_reserved_sms: %d out of range [0, 64]", n);
        return MMB_EINVAL;
    }
    mmb::g_reserved.store(n, std::memory_order_relaxed);
    return MMB_OK;
}
extern "C" long long mmb_launch_count(void) { return mmb::g_launches.load(std::memory_order_relaxed); }
extern "C" int mmb_version(void) { return MMB_VERSION; }
extern "C" const char* mmb_last_error(void) { return mmb::g_err; }
extern "C" int mmb_check_device(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        mmb::set_last_error("no CUDA device");
        return MMB_ECUDA;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        mmb::set_last_error("libmmbert_sm100 needs compute capability 10.x (B200), found %d.x", major);
        return MMB_EARCH;
    }
    return MMB_OK;
}
