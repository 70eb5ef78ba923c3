// libhlv.so -- library-level entry points: version, errors, device info, workspace.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

#include "hlv_common.cuh"

namespace hlv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return HLV_ERR_CUDA;
}

struct DevInfo { int sms, major, minor; };
static DevInfo g_dev[64];
static bool g_dev_ok[64];

static int query_device(DevInfo* out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return HLV_ERR_NO_DEVICE; }
    if (dev < 0 || dev >= 64) { set_error("device index %d out of range", dev); return HLV_ERR_NO_DEVICE; }
    if (!g_dev_ok[dev]) {
        DevInfo d;
        if ((e = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
            (e = cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess ||
            (e = cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess) {
            cuda_fail(e, "cudaDeviceGetAttribute");
            return HLV_ERR_NO_DEVICE;
        }
        g_dev[dev] = d;
        g_dev_ok[dev] = true;
    }
    *out = g_dev[dev];
    return HLV_OK;
}

int sm_count() {
    DevInfo d;
    if (query_device(&d) != HLV_OK) return 0;
    return d.sms;
}

static std::mutex g_cache_mu;
static std::map<std::tuple<int, const void*, int, size_t>, int> g_occupancy;
static std::map<std::pair<int, const void*>, size_t> g_smem_granted;

int cached_occupancy(const void* func, int threads, size_t smem) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 1; }
    const auto key = std::make_tuple(dev, func, threads, smem);
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        auto it = g_occupancy.find(key);
        if (it != g_occupancy.end()) return it->second;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, threads, smem) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        return 1;                                   // not cached: e.g. the shared-memory attribute was not raised yet
    }
    std::lock_guard<std::mutex> lock(g_cache_mu);
    g_occupancy[key] = per_sm;
    return per_sm;
}

cudaError_t ensure_dynamic_smem(const void* func, size_t smem) {
    if (smem <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const auto key = std::make_pair(dev, func);
    {
        std::lock_guard<std::mutex> lock(g_cache_mu);
        auto it = g_smem_granted.find(key);
        if (it != g_smem_granted.end() && it->second >= smem) return cudaSuccess;
    }
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(g_cache_mu);
    size_t& g = g_smem_granted[key];
    if (g < smem) g = smem;
    return cudaSuccess;
}

}  // namespace hlv

extern "C" {

int hlv_version(void) { return HLV_VERSION; }

const char* hlv_last_error_string(void) { return hlv::g_err; }

int hlv_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    hlv::DevInfo d;
    int rc = hlv::query_device(&d);
    if (rc != HLV_OK) return rc;
    if (sm_count) *sm_count = d.sms;
    if (cc_major) *cc_major = d.major;
    if (cc_minor) *cc_minor = d.minor;
    return HLV_OK;
}

size_t hlv_workspace_bytes(int max_rows) { return hlv::workspace_bytes(max_rows); }

int hlv_workspace_init(void* ws, size_t ws_bytes, hlv_stream_t stream) {
    HLV_REQUIRE(ws != nullptr, HLV_ERR_ARG, "hlv_workspace_init: ws is NULL");
    HLV_REQUIRE(ws_bytes >= hlv::workspace_bytes(1), HLV_ERR_WORKSPACE,
                "hlv_workspace_init: ws_bytes=%zu < minimum %zu", ws_bytes, hlv::workspace_bytes(1));
    cudaError_t e = cudaMemsetAsync(ws, 0, hlv::kCounterBytes + hlv::kExtraDoubles * sizeof(double),
                                    static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return hlv::cuda_fail(e, "hlv_workspace_init/cudaMemsetAsync");
    return HLV_OK;
}

}  // extern "C"
