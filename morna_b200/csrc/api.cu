// Library-wide entry points: version, status names, launch counter, device probe.
#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace morna {
thread_local int g_last_cuda_error = 0;
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::mutex g_device_cache_mutex;

int sm_count_current() {
    static std::map<int, int> sms;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_device_cache_mutex);
    int &v = sms[dev];
    if (v <= 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
    }
    return v;
}

int ensure_dynamic_smem(const void *kernel, size_t bytes) {
    static std::map<std::pair<const void *, int>, size_t> granted;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_device_cache_mutex);
    size_t &have = granted[std::make_pair(kernel, dev)];
    if (bytes > have) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return cuda_fail(e);
        have = bytes;
    }
    return MORNA_OK;
}
}  // namespace morna

extern "C" int morna_abi_version(void) { return MORNA_ABI_VERSION; }

extern "C" const char *morna_status_string(int status) {
    switch (status) {
        case MORNA_OK: return "ok";
        case MORNA_ERR_INVALID_ARGUMENT: return "invalid argument";
        case MORNA_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case MORNA_ERR_CUDA: return "CUDA call failed";
        case MORNA_ERR_UNSUPPORTED_DEVICE: return "unsupported device (needs sm_100)";
        case MORNA_ERR_NO_SAMPLES: return "no internal ids were assigned";
        case MORNA_ERR_CAPACITY: return "output capacity exceeded";
        default: return "unknown status";
    }
}

extern "C" int morna_last_cuda_error(void) { return morna::g_last_cuda_error; }

extern "C" int64_t morna_kernel_launch_count(void) { return (int64_t)morna::g_launches.load(); }

extern "C" int morna_note_graph_replay(int64_t kernels) {
    if (kernels < 0) return MORNA_ERR_INVALID_ARGUMENT;
    morna::g_launches.fetch_add(kernels, std::memory_order_relaxed);
    return MORNA_OK;
}

extern "C" int morna_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor) {
    int dev = 0, v = 0;
    MORNA_CUDA_TRY(cudaGetDevice(&dev));
    if (sm_count) { MORNA_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (cc_major) { MORNA_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { MORNA_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    return MORNA_OK;
}
