// libhlv.so -- the exchange steps of the sharded recurrence over NVLink peer memory (see hlv_peer.cuh):
// exchange-area management, flag-only signal / wait, and the reduce-scatter + alpha kernel.
#include <stdlib.h>

#include "hlv_peer.cuh"

namespace hlv {

__global__ void peer_signal_kernel(const __grid_constant__ PeerView pv, int channel) {
    __threadfence_system();                      // everything earlier kernels of this stream wrote is published with the flag
    __syncthreads();
    peer_push(pv, channel, nullptr, 0);
}

__global__ void peer_wait_kernel(const __grid_constant__ PeerView pv, int channel) { (void)peer_wait_all(pv, channel); }

// ---- reduce-scatter + alpha ---------------------------------------------------------------------------------
// w[i] = hv_0[lo+i] + hv_1[lo+i] + ... (rank order), alpha partial = sum w[i] v[i].  One 128-bit load per rank and
// chunk in flight per thread (world independent loads: the NVLink round trip is ~2 us, so the memory-level
// parallelism comes from the ranks, the unroll and the resident CTAs), peer loads are plain weak loads issued after the
// system-scope acquire of the HLV_CH_HV flags.
struct HvTable {
    const float* hv[HLV_MAX_PEERS];
};
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}

// In-switch reduction (NVLS): one load of the multicast address returns the sum over every rank's copy.
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}

template <int WORLD>   // 0 = run-time world; -1 = multicast (the switch adds)
__global__ void __launch_bounds__(kThreads)
reduce_scatter_dot_kernel(const __grid_constant__ PeerView pv, const __grid_constant__ HvTable tab, int64_t lo, int64_t n,
                          float* __restrict__ w, const float* __restrict__ v, double* partials, unsigned* counter,
                          double* alpha_out) {
    __shared__ double s_warp[kWarps];
    const int world = WORLD > 0 ? WORLD : pv.world;
    if (pv.world > 1) (void)peer_wait_all(pv, HLV_CH_HV);
    const int64_t nvec = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    float acc[2] = {0.f, 0.f};
    auto one = [&](int64_t i, float& a) {
        float4 s;
        if constexpr (WORLD < 0) {
            s = multimem_ld_reduce_f4(tab.hv[HLV_MAX_PEERS - 1] + lo + 4 * i);     // the multicast address rides in the last slot
        } else {
            constexpr int NP = WORLD ? WORLD : HLV_MAX_PEERS;
            float4 x[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p)
                if (p < world) x[p] = ld_peer_f4(tab.hv[p] + lo + 4 * i);
            s = x[0];
#pragma unroll
            for (int p = 1; p < NP; ++p)
                if (p < world) { s.x += x[p].x; s.y += x[p].y; s.z += x[p].z; s.w += x[p].w; }
        }
        const float4 y = ldg_stream(reinterpret_cast<const float4*>(v) + i);
        reinterpret_cast<float4*>(w)[i] = s;
        a = fmaf(s.x, y.x, a); a = fmaf(s.y, y.y, a); a = fmaf(s.z, y.z, a); a = fmaf(s.w, y.w, a);
    };
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + stride < nvec; i += 2 * stride) { one(i, acc[0]); one(i + stride, acc[1]); }
    for (; i < nvec; i += stride) one(i, acc[0]);
    const int64_t t = (nvec << 2) + (int64_t)blockIdx.x * kThreads + threadIdx.x;      // ragged tail (< 4 elements)
    if (t < n) {
        float s = tab.hv[0][lo + t];
        for (int p = 1; p < world; ++p) s += tab.hv[p][lo + t];
        w[t] = s;
        acc[1] = fmaf(s, v[t], acc[1]);
    }
    double s = block_sum((double)acc[0] + (double)acc[1], s_warp);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
    finalize_rows_push(partials, counter, 1, alpha_out, pv, HLV_CH_ALPHA);
}

}  // namespace hlv

using namespace hlv;

extern "C" {

size_t hlv_peer_xchg_bytes(void) { return kXchgBytes; }

int hlv_peer_xchg_init(void* xchg_local, hlv_stream_t stream) {
    HLV_REQUIRE(xchg_local != nullptr && aligned16(xchg_local), HLV_ERR_ARG, "hlv_peer_xchg_init: NULL or unaligned exchange area");
    cudaError_t e = cudaMemsetAsync(xchg_local, 0, kXchgBytes, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "hlv_peer_xchg_init/cudaMemsetAsync");
    return HLV_OK;
}

int hlv_peer_xchg_error(const void* xchg_local, int* h_error_out, hlv_stream_t stream) {
    HLV_REQUIRE(xchg_local && h_error_out, HLV_ERR_ARG, "hlv_peer_xchg_error: bad argument");
    unsigned v = 0;
    cudaError_t e = cudaMemcpyAsync(&v, static_cast<const char*>(xchg_local) + kXchgErrorOff, sizeof(v), cudaMemcpyDeviceToHost,
                                    static_cast<cudaStream_t>(stream));
    if (e == cudaSuccess) e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "hlv_peer_xchg_error");
    *h_error_out = (int)v;
    return HLV_OK;
}

int hlv_peer_signal(const hlv_peer_ctx* h_ctx, int channel, hlv_stream_t stream) {
    int rc = check_peer_ctx(h_ctx, "hlv_peer_signal");
    if (rc != HLV_OK) return rc;
    HLV_REQUIRE(channel >= 0 && channel < kChannels, HLV_ERR_ARG, "hlv_peer_signal: channel %d out of range", channel);
    if (h_ctx == nullptr || h_ctx->world <= 1) return HLV_OK;
    peer_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(make_peer_view(h_ctx), channel);
    HLV_LAUNCH_CHECK("hlv_peer_signal");
    return HLV_OK;
}

int hlv_peer_wait(const hlv_peer_ctx* h_ctx, int channel, hlv_stream_t stream) {
    int rc = check_peer_ctx(h_ctx, "hlv_peer_wait");
    if (rc != HLV_OK) return rc;
    HLV_REQUIRE(channel >= 0 && channel < kChannels, HLV_ERR_ARG, "hlv_peer_wait: channel %d out of range", channel);
    if (h_ctx == nullptr || h_ctx->world <= 1) return HLV_OK;
    peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(make_peer_view(h_ctx), channel);
    HLV_LAUNCH_CHECK("hlv_peer_wait");
    return HLV_OK;
}

int hlv_x_reduce_scatter_dot_f32(const hlv_peer_ctx* h_ctx, const float* const* h_hv, const float* hv_multicast,
                                 int64_t shard_lo, int64_t n, float* w, const float* v, double* alpha_out,
                                 void* ws_raw, size_t ws_bytes, hlv_stream_t stream) {
    int rc = check_peer_ctx(h_ctx, "hlv_x_reduce_scatter_dot_f32");
    if (rc != HLV_OK) return rc;
    const int world = h_ctx ? h_ctx->world : 1;
    HLV_REQUIRE(h_hv && w && v && alpha_out && n >= 0 && shard_lo >= 0, HLV_ERR_ARG, "hlv_x_reduce_scatter_dot_f32: bad argument");
    HLV_REQUIRE((shard_lo & 3) == 0 && aligned16(w) && aligned16(v), HLV_ERR_ALIGN,
                "hlv_x_reduce_scatter_dot_f32: w, v must be 16-byte aligned and shard_lo a multiple of 4");
    HvTable tab{};
    for (int p = 0; p < world; ++p) {
        HLV_REQUIRE(h_hv[p] != nullptr && aligned16(h_hv[p]), HLV_ERR_ALIGN, "hlv_x_reduce_scatter_dot_f32: Hv of rank %d NULL or unaligned", p);
        tab.hv[p] = h_hv[p];
    }
    Workspace ws;
    HLV_REQUIRE(carve_workspace(ws_raw, ws_bytes, 1, &ws), HLV_ERR_WORKSPACE, "hlv_x_reduce_scatter_dot_f32: workspace too small");
    HLV_REQUIRE(sm_count() > 0, HLV_ERR_NO_DEVICE, "hlv_x_reduce_scatter_dot_f32: no CUDA device");
    const PeerView pv = make_peer_view(h_ctx);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t items = (n / 4 + kThreads * 2 - 1) / (kThreads * 2) + 1;
#define HLV_RS_LAUNCH(W)                                                                                              \
    do {                                                                                                              \
        const int grid = persistent_grid(items, cached_resident_ctas(reduce_scatter_dot_kernel<W>, kThreads, 0));      \
        reduce_scatter_dot_kernel<W><<<grid, kThreads, 0, s>>>(pv, tab, shard_lo, n, w, v, ws.partials, ws.counters, alpha_out); \
    } while (0)
    const bool mc = hv_multicast != nullptr && world > 1 && world < HLV_MAX_PEERS && aligned16(hv_multicast) && (n & 3) == 0;
    if (mc) tab.hv[HLV_MAX_PEERS - 1] = hv_multicast;
    if (mc) { HLV_RS_LAUNCH(-1); } else
    switch (world) {
        case 1: HLV_RS_LAUNCH(1); break;
        case 2: HLV_RS_LAUNCH(2); break;
        case 4: HLV_RS_LAUNCH(4); break;
        case 8: HLV_RS_LAUNCH(8); break;
        default: HLV_RS_LAUNCH(0); break;
    }
#undef HLV_RS_LAUNCH
    HLV_LAUNCH_CHECK("hlv_x_reduce_scatter_dot_f32");
    return HLV_OK;
}

}  // extern "C"
