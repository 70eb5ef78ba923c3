// Shared device/host helpers for libhlv.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "hlv.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libhlv is written for sm_100a (B200) only"
#endif

namespace hlv {

constexpr int kThreads = 256;            // every streaming kernel: 8 warps / CTA
constexpr int kWarps = kThreads / 32;
constexpr int kMaxCtas = 2048;           // upper bound on any persistent grid (partials stride)
constexpr int kCounterBytes = 256;       // ws head: ticket counters
constexpr int kExtraDoubles = 32;        // ws: small scratch after the counters

// ---- workspace layout -------------------------------------------------------
//   [0,256)            unsigned ticket counters (self-resetting, see finalize_rows)
//   [256,512)          kExtraDoubles doubles of scratch
//   [512, ...)         partials[row][kMaxCtas] doubles (row-major by row => the final
//                      stage reads one row with coalesced loads)
struct Workspace {
    unsigned* counters;
    double* extra;
    double* partials;
    int max_rows;
};

inline size_t workspace_bytes(int max_rows) {
    if (max_rows < 1) max_rows = 1;
    return (size_t)kCounterBytes + kExtraDoubles * sizeof(double) +
           (size_t)max_rows * kMaxCtas * sizeof(double);
}

inline bool carve_workspace(void* ws, size_t ws_bytes, int rows, Workspace* out) {
    if (ws == nullptr || ws_bytes < workspace_bytes(rows)) return false;
    char* p = static_cast<char*>(ws);
    out->counters = reinterpret_cast<unsigned*>(p);
    out->extra = reinterpret_cast<double*>(p + kCounterBytes);
    out->partials = reinterpret_cast<double*>(p + kCounterBytes + kExtraDoubles * sizeof(double));
    out->max_rows = (int)((ws_bytes - kCounterBytes - kExtraDoubles * sizeof(double)) /
                          (kMaxCtas * sizeof(double)));
    return true;
}

// ---- host-side error plumbing (api.cu) ---------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();                                  // cached per device, <=0 on failure

#define HLV_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            ::hlv::set_error(__VA_ARGS__);      \
            return (code);                      \
        }                                       \
    } while (0)

#define HLV_LAUNCH_CHECK(what)                                   \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return ::hlv::cuda_fail(e__, what); \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Resident CTAs per SM of a kernel from the occupancy API, so a persistent grid is exactly ONE resident wave: no
// second, under-filled wave and no tail.  The answer is cached per (device, kernel, threads, shared memory): the
// query costs microseconds on the host and the recurrence issues ~6 launches per iteration.
int cached_occupancy(const void* func, int threads, size_t smem);                      // hlv_api.cu
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) only when `smem` exceeds what was already granted on this device
cudaError_t ensure_dynamic_smem(const void* func, size_t smem);                        // hlv_api.cu
template <typename Kernel>
inline int cached_resident_ctas(Kernel kernel, int threads, size_t smem) {
    return cached_occupancy(reinterpret_cast<const void*>(kernel), threads, smem);
}
template <typename Kernel>
inline int resident_ctas(Kernel kernel, size_t smem = 0) {
    return cached_occupancy(reinterpret_cast<const void*>(kernel), kThreads, smem);
}

// Persistent grid: one resident wave, capped by the amount of work.
inline int persistent_grid(int64_t work_items, int ctas_per_sm) {
    int64_t g = (int64_t)sm_count() * ctas_per_sm;
    if (g > kMaxCtas) g = kMaxCtas;
    if (g > work_items) g = work_items;
    if (g < 1) g = 1;
    return (int)g;
}

#ifdef __CUDACC__
// ---- streaming loads: read-once data bypasses L1 allocation -------------------
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// ---- warp / block reductions (fixed order => deterministic) --------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduce 8 per-lane values across the warp with 9 shuffles instead of 40:
// each butterfly step halves the number of live values.  On return, lane l with
// (l & 3) == 0 holds the warp total of value index ((l>>4)&1)*4 + ((l>>3)&1)*2 + ((l>>2)&1).
__device__ __forceinline__ float warp_sum8(float (&v)[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float keep = b4 ? v[i + 4] : v[i];
        float send = b4 ? v[i] : v[i + 4];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float b[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float keep = b3 ? a[i + 2] : a[i];
        float send = b3 ? a[i] : a[i + 2];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float keep = b2 ? b[1] : b[0];
    float send = b2 ? b[0] : b[1];
    float c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    return c;
}
__device__ __forceinline__ int warp_sum8_row(int lane) {
    return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

// Block-wide sum of one double per thread; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* s_warp /* [kWarps] */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < kWarps; ++i) t += s_warp[i];
    }
    return t;
}

// Cross-CTA final stage.  Precondition: this CTA has written its per-row partials to
// partials[r*kMaxCtas + blockIdx.x].  The last CTA to arrive (ticket counter, wraps to 0
// by itself so the workspace needs zeroing only once) sums every row over CTAs in index
// order: lane-strided fp64 accumulation + xor butterfly -- independent of arrival order.
__device__ __forceinline__ void finalize_rows(const double* partials, unsigned* counter,
                                              int rows, double* out) {
    __shared__ unsigned s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicInc(counter, gridDim.x - 1);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nblk = gridDim.x;
    for (int r = warp; r < rows; r += kWarps) {
        const double* p = partials + (size_t)r * kMaxCtas;
        double s = 0.0;
        for (int b = lane; b < nblk; b += 32) s += __ldcg(p + b);
        s = warp_sum(s);
        if (lane == 0) out[r] = s;
    }
}
#endif  // __CUDACC__

}  // namespace hlv
