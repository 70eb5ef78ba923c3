// libhlv.so -- kernels (a) multi-tensor gather/scatter and (b) the Lanczos recurrence:
// fused dot, fused three-term update + norm, normalise + store.
//
// All of these are pure HBM streams (<= 0.5 flop/byte): 128-bit coalesced accesses,
// several independent loads in flight per thread, persistent grids sized from the SM count,
// fixed-order reductions with an fp64 final stage (hlv_common.cuh).
#include <stdlib.h>

#include "hlv_peer.cuh"

namespace hlv {

// =============================================================================
// (a) gather / scatter
// =============================================================================
// The per-call pointer table travels as a kernel parameter (CUDA >= 12.1 allows 32 KB of
// parameters), so a call needs no staging copy, no hidden allocation and no sync, and the
// launch can be captured in a CUDA graph.  start[t] is tensor t's offset in the flat vector.
template <int CAP>
struct TensorTable {
    const void* ptr[CAP];
    int64_t start[CAP + 1];
    int count;
};

constexpr int kChunk = 8192;            // flat elements per CTA work item (32 KB)
constexpr int kSmallTable = 224;        // table that fits the classic 4 KB parameter space

enum GatherMode { kCopy = 0, kScaleAcc = 1 };

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// One contiguous segment: flat[0..len) <-> tens[0..len).  GATHER: tens -> flat (+ dot with v).
template <bool GATHER, int MODE, bool DOT>
__device__ __forceinline__ void move_segment(float* __restrict__ flat, float* __restrict__ tens,
                                             const float* __restrict__ v, int64_t len, float scale,
                                             bool accumulate, float& dot) {
    const int tid = threadIdx.x;
    const float* src = GATHER ? tens : flat;
    float* dst = GATHER ? flat : tens;
    auto body1 = [&](int64_t i) {
        float x = src[i];
        if (MODE == kScaleAcc) x = (accumulate ? dst[i] : 0.0f) + scale * x;
        dst[i] = x;
        if (DOT) dot = fmaf(x, v[i], dot);
    };
    const uintptr_t a_src = reinterpret_cast<uintptr_t>(src), a_dst = reinterpret_cast<uintptr_t>(dst);
    if (((a_src ^ a_dst) & 15u) != 0) {                 // relative misalignment: scalar, still coalesced
        for (int64_t i = tid; i < len; i += kThreads) body1(i);
        return;
    }
    int64_t head = ((16 - (a_dst & 15u)) & 15u) >> 2;   // elements until 16-byte boundary
    if (head > len) head = len;
    if (tid < head) body1(tid);
    const int64_t nvec = (len - head) >> 2;
    const float* s4 = src + head;
    float* d4 = dst + head;
    const float* v4 = DOT ? v + head : nullptr;
    int64_t i = tid;
    // 4 independent 128-bit loads in flight per thread
    for (; i + 3 * kThreads < nvec; i += 4 * kThreads) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = ld4(s4 + 4 * (i + u * kThreads));
        if (MODE == kScaleAcc) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float4 o = accumulate ? ld4(d4 + 4 * (i + u * kThreads)) : make_float4(0.f, 0.f, 0.f, 0.f);
                x[u].x = o.x + scale * x[u].x; x[u].y = o.y + scale * x[u].y;
                x[u].z = o.z + scale * x[u].z; x[u].w = o.w + scale * x[u].w;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) st4(d4 + 4 * (i + u * kThreads), x[u]);
        if (DOT) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float4 y = ld4(v4 + 4 * (i + u * kThreads));
                dot = fmaf(x[u].x, y.x, dot); dot = fmaf(x[u].y, y.y, dot);
                dot = fmaf(x[u].z, y.z, dot); dot = fmaf(x[u].w, y.w, dot);
            }
        }
    }
    for (; i < nvec; i += kThreads) {
        float4 x = ld4(s4 + 4 * i);
        if (MODE == kScaleAcc) {
            float4 o = accumulate ? ld4(d4 + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            x.x = o.x + scale * x.x; x.y = o.y + scale * x.y; x.z = o.z + scale * x.z; x.w = o.w + scale * x.w;
        }
        st4(d4 + 4 * i, x);
        if (DOT) {
            float4 y = ld4(v4 + 4 * i);
            dot = fmaf(x.x, y.x, dot); dot = fmaf(x.y, y.y, dot); dot = fmaf(x.z, y.z, dot); dot = fmaf(x.w, y.w, dot);
        }
    }
    const int64_t tail0 = head + (nvec << 2);
    if (tail0 + tid < len) body1(tail0 + tid);
}

template <int CAP, bool GATHER, int MODE, bool DOT>
__global__ void __launch_bounds__(kThreads)
multi_tensor_kernel(const __grid_constant__ TensorTable<CAP> tab, float* __restrict__ flat,
                    const float* __restrict__ v, float scale, int accumulate,
                    double* partials, unsigned* counter, double* dot_out) {
    __shared__ double s_warp[kWarps];
    const int64_t lo = tab.start[0], hi = tab.start[tab.count];
    const int64_t nchunks = (hi - lo + kChunk - 1) / kChunk;
    float dot = 0.0f;
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        int64_t pos = lo + c * kChunk;
        const int64_t end = (pos + kChunk < hi) ? pos + kChunk : hi;
        // last tensor with start <= pos (uniform across the CTA: constant-bank reads broadcast)
        int a = 0, b = tab.count - 1;
        while (a < b) {
            int m = (a + b + 1) >> 1;
            if (tab.start[m] <= pos) a = m; else b = m - 1;
        }
        int t = a;
        while (pos < end) {
            while (tab.start[t + 1] <= pos) ++t;        // skips empty tensors
            const int64_t seg_end = tab.start[t + 1] < end ? tab.start[t + 1] : end;
            float* tens = const_cast<float*>(static_cast<const float*>(tab.ptr[t])) + (pos - tab.start[t]);
            move_segment<GATHER, MODE, DOT>(flat + pos, tens, DOT ? v + pos : nullptr, seg_end - pos,
                                            scale, accumulate != 0, dot);
            pos = seg_end;
        }
    }
    if (DOT) {
        double t = block_sum((double)dot, s_warp);
        if (threadIdx.x == 0) partials[blockIdx.x] = t;
        finalize_rows(partials, counter, 1, dot_out);
    }
}

__global__ void sum_small_kernel(const double* in, int count, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < count; ++i) s += in[i];
        out[0] = s;
    }
}

template <int CAP, bool GATHER>
static int launch_multi_tensor(const void* const* h_ptr, const int64_t* h_numel, int t0, int t1,
                               int64_t flat_off, float* flat, const float* v, float scale, int accumulate,
                               const Workspace* ws, double* dot_out, cudaStream_t stream) {
    TensorTable<CAP> tab;
    tab.count = t1 - t0;
    int64_t off = flat_off;
    for (int t = t0; t < t1; ++t) {
        tab.ptr[t - t0] = h_ptr[t];
        tab.start[t - t0] = off;
        off += h_numel[t];
    }
    tab.start[tab.count] = off;
    const int64_t total = off - flat_off;
    if (total == 0) {
        if (dot_out) {
            cudaError_t e = cudaMemsetAsync(dot_out, 0, sizeof(double), stream);
            if (e != cudaSuccess) return cuda_fail(e, "gather/memset");
        }
        return HLV_OK;
    }
    const bool plain = (scale == 1.0f && !accumulate);
    int per_sm;
    if (!GATHER) per_sm = resident_ctas(multi_tensor_kernel<CAP, false, kCopy, false>);
    else if (dot_out) per_sm = plain ? resident_ctas(multi_tensor_kernel<CAP, true, kCopy, true>)
                                     : resident_ctas(multi_tensor_kernel<CAP, true, kScaleAcc, true>);
    else per_sm = plain ? resident_ctas(multi_tensor_kernel<CAP, true, kCopy, false>)
                        : resident_ctas(multi_tensor_kernel<CAP, true, kScaleAcc, false>);
    const int grid = persistent_grid((total + kChunk - 1) / kChunk, per_sm);
    if (!GATHER) {
        multi_tensor_kernel<CAP, false, kCopy, false><<<grid, kThreads, 0, stream>>>(
            tab, flat, nullptr, 1.0f, 0, nullptr, nullptr, nullptr);
    } else if (dot_out) {
        if (plain)
            multi_tensor_kernel<CAP, true, kCopy, true><<<grid, kThreads, 0, stream>>>(
                tab, flat, v, scale, accumulate, ws->partials, ws->counters, dot_out);
        else
            multi_tensor_kernel<CAP, true, kScaleAcc, true><<<grid, kThreads, 0, stream>>>(
                tab, flat, v, scale, accumulate, ws->partials, ws->counters, dot_out);
    } else {
        if (plain)
            multi_tensor_kernel<CAP, true, kCopy, false><<<grid, kThreads, 0, stream>>>(
                tab, flat, nullptr, scale, accumulate, nullptr, nullptr, nullptr);
        else
            multi_tensor_kernel<CAP, true, kScaleAcc, false><<<grid, kThreads, 0, stream>>>(
                tab, flat, nullptr, scale, accumulate, nullptr, nullptr, nullptr);
    }
    HLV_LAUNCH_CHECK(GATHER ? "hlv_gather_f32 launch" : "hlv_scatter_f32 launch");
    return HLV_OK;
}

template <bool GATHER>
static int multi_tensor(const void* const* h_ptr, const int64_t* h_numel, int ntensors, float* flat,
                        int64_t flat_len, float scale, int accumulate, const float* v, double* dot_out,
                        void* ws_raw, size_t ws_bytes, cudaStream_t stream, const char* name) {
    HLV_REQUIRE(ntensors >= 0 && (ntensors == 0 || (h_ptr && h_numel)), HLV_ERR_ARG, "%s: bad tensor list", name);
    HLV_REQUIRE(flat != nullptr || flat_len == 0, HLV_ERR_ARG, "%s: flat vector is NULL", name);
    HLV_REQUIRE((v == nullptr) == (dot_out == nullptr), HLV_ERR_ARG, "%s: v and dot_out must both be set or both NULL", name);
    int64_t total = 0;
    for (int t = 0; t < ntensors; ++t) {
        HLV_REQUIRE(h_numel[t] >= 0, HLV_ERR_ARG, "%s: numel[%d] < 0", name, t);
        HLV_REQUIRE(h_numel[t] == 0 || h_ptr[t] != nullptr, HLV_ERR_ARG, "%s: tensor %d is NULL", name, t);
        HLV_REQUIRE((reinterpret_cast<uintptr_t>(h_ptr[t]) & 3u) == 0, HLV_ERR_ALIGN, "%s: tensor %d not 4-byte aligned", name, t);
        total += h_numel[t];
    }
    HLV_REQUIRE(total == flat_len, HLV_ERR_ARG, "%s: sum(numel)=%lld != flat length %lld", name,
                (long long)total, (long long)flat_len);
    HLV_REQUIRE((reinterpret_cast<uintptr_t>(flat) & 3u) == 0, HLV_ERR_ALIGN, "%s: flat vector not 4-byte aligned", name);
    Workspace ws{};
    if (dot_out) {
        HLV_REQUIRE(carve_workspace(ws_raw, ws_bytes, 1, &ws), HLV_ERR_WORKSPACE, "%s: workspace too small", name);
        HLV_REQUIRE(((reinterpret_cast<uintptr_t>(flat) ^ reinterpret_cast<uintptr_t>(v)) & 15u) == 0, HLV_ERR_ALIGN,
                    "%s: v and dst must share 16-byte alignment phase", name);
    }
    HLV_REQUIRE(sm_count() > 0, HLV_ERR_NO_DEVICE, "%s: no CUDA device", name);
    if (ntensors <= kSmallTable)
        return launch_multi_tensor<kSmallTable, GATHER>(h_ptr, h_numel, 0, ntensors, 0, flat, v, scale, accumulate,
                                                &ws, dot_out, stream);
    // Long lists: chunks of HLV_MAX_TENSORS; per-chunk dots land in ws.extra and are summed in order.
    const int nchunk = (ntensors + HLV_MAX_TENSORS - 1) / HLV_MAX_TENSORS;
    HLV_REQUIRE(!dot_out || nchunk <= kExtraDoubles, HLV_ERR_ARG, "%s: too many tensors (%d) for fused dot", name, ntensors);
    int64_t off = 0;
    for (int c = 0; c < nchunk; ++c) {
        const int t0 = c * HLV_MAX_TENSORS, t1 = (t0 + HLV_MAX_TENSORS < ntensors) ? t0 + HLV_MAX_TENSORS : ntensors;
        double* out_c = dot_out ? (nchunk == 1 ? dot_out : ws.extra + c) : nullptr;
        int rc = launch_multi_tensor<HLV_MAX_TENSORS, GATHER>(h_ptr, h_numel, t0, t1, off, flat, v, scale,
                                                              accumulate, &ws, out_c, stream);
        if (rc != HLV_OK) return rc;
        for (int t = t0; t < t1; ++t) off += h_numel[t];
    }
    if (dot_out && nchunk > 1) {
        sum_small_kernel<<<1, 32, 0, stream>>>(ws.extra, nchunk, dot_out);
        HLV_LAUNCH_CHECK("gather/sum_small");
    }
    return HLV_OK;
}

// =============================================================================
// (b) recurrence
// =============================================================================
constexpr int kVecPerThread = 4;        // independent 128-bit loads in flight per thread and operand

__global__ void __launch_bounds__(kThreads)
dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
           double* partials, unsigned* counter, double* out) {
    __shared__ double s_warp[kWarps];
    const int64_t nvec = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float acc[kVecPerThread] = {0.f, 0.f, 0.f, 0.f};
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + (kVecPerThread - 1) * stride < nvec; i += kVecPerThread * stride) {
        float4 x[kVecPerThread], y[kVecPerThread];
#pragma unroll
        for (int u = 0; u < kVecPerThread; ++u) { x[u] = ldg_stream(a4 + i + u * stride); y[u] = ldg_stream(b4 + i + u * stride); }
#pragma unroll
        for (int u = 0; u < kVecPerThread; ++u) {
            acc[u] = fmaf(x[u].x, y[u].x, acc[u]); acc[u] = fmaf(x[u].y, y[u].y, acc[u]);
            acc[u] = fmaf(x[u].z, y[u].z, acc[u]); acc[u] = fmaf(x[u].w, y[u].w, acc[u]);
        }
    }
    for (; i < nvec; i += stride) {
        float4 x = ldg_stream(a4 + i), y = ldg_stream(b4 + i);
        acc[0] = fmaf(x.x, y.x, acc[0]); acc[0] = fmaf(x.y, y.y, acc[0]);
        acc[0] = fmaf(x.z, y.z, acc[0]); acc[0] = fmaf(x.w, y.w, acc[0]);
    }
    const int64_t t = (nvec << 2) + (int64_t)blockIdx.x * kThreads + threadIdx.x;   // ragged tail (< 4 elements)
    if (t < n) acc[1] = fmaf(a[t], b[t], acc[1]);
    double s = ((double)acc[0] + (double)acc[1]) + ((double)acc[2] + (double)acc[3]);
    s = block_sum(s, s_warp);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
    finalize_rows(partials, counter, 1, out);
}

// w -= alpha*v + beta*v_old, with the reference's rounding sequence (two products, one sum,
// one subtraction -- no FMA contraction) so that, given the same alpha/beta, the result is
// bit-identical to torch's elementwise ops (lanczostrain_hand.py:185,202); then sum w^2.
template <bool HAS_OLD>
__global__ void __launch_bounds__(kThreads)
lanczos_update_kernel(float* __restrict__ w, const float* __restrict__ vj, const float* __restrict__ vo,
                      double* alpha_p, const double* __restrict__ beta_p, int64_t n,
                      double* partials, unsigned* counter, double* norm2_out, const __grid_constant__ PeerView pv) {
    __shared__ double s_warp[kWarps];
    double alpha_d;
    if (pv.world > 1) {                                   // alpha = rank-ordered total of the partial dot products
        alpha_d = peer_pull_scalar(pv, HLV_CH_ALPHA);
        if (blockIdx.x == 0 && threadIdx.x == 0) alpha_p[0] = alpha_d;
    } else {
        alpha_d = alpha_p[0];
    }
    const float alpha = (float)alpha_d;
    const float beta = HAS_OLD ? (float)beta_p[0] : 0.0f;
    auto f = [&](float wv, float a, float b) -> float {
        float t = __fmul_rn(alpha, a);
        if (HAS_OLD) t = __fadd_rn(t, __fmul_rn(beta, b));
        return __fsub_rn(wv, t);
    };
    const int64_t nvec = n >> 2;
    float4* w4 = reinterpret_cast<float4*>(w);
    const float4* a4 = reinterpret_cast<const float4*>(vj);
    const float4* b4 = reinterpret_cast<const float4*>(vo);
    float acc[kVecPerThread] = {0.f, 0.f, 0.f, 0.f};
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + (kVecPerThread - 1) * stride < nvec; i += kVecPerThread * stride) {
        float4 x[kVecPerThread], a[kVecPerThread], b[kVecPerThread];
#pragma unroll
        for (int u = 0; u < kVecPerThread; ++u) {
            x[u] = w4[i + u * stride];
            a[u] = ldg_stream(a4 + i + u * stride);
            if (HAS_OLD) b[u] = ldg_stream(b4 + i + u * stride);
            else b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kVecPerThread; ++u) {
            x[u].x = f(x[u].x, a[u].x, b[u].x); x[u].y = f(x[u].y, a[u].y, b[u].y);
            x[u].z = f(x[u].z, a[u].z, b[u].z); x[u].w = f(x[u].w, a[u].w, b[u].w);
            w4[i + u * stride] = x[u];
            acc[u] = fmaf(x[u].x, x[u].x, acc[u]); acc[u] = fmaf(x[u].y, x[u].y, acc[u]);
            acc[u] = fmaf(x[u].z, x[u].z, acc[u]); acc[u] = fmaf(x[u].w, x[u].w, acc[u]);
        }
    }
    for (; i < nvec; i += stride) {
        float4 x = w4[i], a = ldg_stream(a4 + i);
        float4 b = HAS_OLD ? ldg_stream(b4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        x.x = f(x.x, a.x, b.x); x.y = f(x.y, a.y, b.y); x.z = f(x.z, a.z, b.z); x.w = f(x.w, a.w, b.w);
        w4[i] = x;
        acc[0] = fmaf(x.x, x.x, acc[0]); acc[0] = fmaf(x.y, x.y, acc[0]);
        acc[0] = fmaf(x.z, x.z, acc[0]); acc[0] = fmaf(x.w, x.w, acc[0]);
    }
    const int64_t t = (nvec << 2) + (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t < n) {
        float x = f(w[t], vj[t], HAS_OLD ? vo[t] : 0.0f);
        w[t] = x;
        acc[1] = fmaf(x, x, acc[1]);
    }
    double s = ((double)acc[0] + (double)acc[1]) + ((double)acc[2] + (double)acc[3]);
    s = block_sum(s, s_warp);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
    finalize_rows_push(partials, counter, 1, norm2_out, pv, HLV_CH_NORM);
}

// beta = sqrt(norm2); v = w / beta (true division, as torch does); optional bf16 row copy.
// With a peer view: norm2 = rank-ordered total of HLV_CH_NORM, and the normalised shard is ALSO stored into every
// rank's full-length vector at shard_lo (plain peer stores: the all-gather of v_{j+1} without a collective launch);
// when every CTA's stores are fenced at system scope the last CTA raises HLV_CH_V on every rank.
struct VecTable {
    float* v[HLV_MAX_PEERS];
    float* mc;                      // NVSwitch multicast address of the same buffers, or NULL
    int count;
};
__device__ __forceinline__ void multimem_st_f4(float* mc, float4 x) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
}
template <bool BF16, bool PEER>
__global__ void __launch_bounds__(kThreads)
normalize_store_kernel(const float* __restrict__ w, double* norm2, int64_t n,
                       double* beta_out, float* __restrict__ v_out, uint16_t* __restrict__ row_bf16,
                       double breakdown_tol, int* breakdown_iter, int iter,
                       const __grid_constant__ PeerView pv, const __grid_constant__ VecTable vt, int64_t shard_lo, unsigned* counter) {
    double n2;
    if (PEER && pv.world > 1) {
        n2 = peer_pull_scalar(pv, HLV_CH_NORM);
        if (blockIdx.x == 0 && threadIdx.x == 0) norm2[0] = n2;
    } else {
        n2 = norm2[0];
    }
    const double beta_d = sqrt(n2);
    const float beta = (float)beta_d;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        beta_out[0] = beta_d;
        if (breakdown_iter != nullptr && beta_d < breakdown_tol && *breakdown_iter < 0) *breakdown_iter = iter;
    }
    const int64_t n8 = n >> 3;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    auto emit = [&](int64_t i, float4 x0, float4 x1) {
        x0.x = __fdiv_rn(x0.x, beta); x0.y = __fdiv_rn(x0.y, beta); x0.z = __fdiv_rn(x0.z, beta); x0.w = __fdiv_rn(x0.w, beta);
        x1.x = __fdiv_rn(x1.x, beta); x1.y = __fdiv_rn(x1.y, beta); x1.z = __fdiv_rn(x1.z, beta); x1.w = __fdiv_rn(x1.w, beta);
        if (v_out != nullptr) {
            reinterpret_cast<float4*>(v_out)[2 * i] = x0;
            reinterpret_cast<float4*>(v_out)[2 * i + 1] = x1;
        }
        if (PEER) {
            if (vt.mc != nullptr) {                         // one store, replicated to every rank by the switch
                multimem_st_f4(vt.mc + shard_lo + 8 * i, x0);
                multimem_st_f4(vt.mc + shard_lo + 8 * i + 4, x1);
            } else {
#pragma unroll 4
                for (int p = 0; p < vt.count; ++p) {
                    float4* dst = reinterpret_cast<float4*>(vt.v[p] + shard_lo);
                    dst[2 * i] = x0;
                    dst[2 * i + 1] = x1;
                }
            }
        }
        if (BF16) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(x0.x, x0.y), p1 = __floats2bfloat162_rn(x0.z, x0.w);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(x1.x, x1.y), p3 = __floats2bfloat162_rn(x1.z, x1.w);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
            o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
            reinterpret_cast<uint4*>(row_bf16)[i] = o;
        }
    };
    const float4* w4 = reinterpret_cast<const float4*>(w);
    int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    for (; i + stride < n8; i += 2 * stride) {             // 4 independent 128-bit loads in flight per thread
        float4 a0 = ldg_stream(w4 + 2 * i), a1 = ldg_stream(w4 + 2 * i + 1);
        float4 b0 = ldg_stream(w4 + 2 * (i + stride)), b1 = ldg_stream(w4 + 2 * (i + stride) + 1);
        emit(i, a0, a1);
        emit(i + stride, b0, b1);
    }
    for (; i < n8; i += stride) emit(i, ldg_stream(w4 + 2 * i), ldg_stream(w4 + 2 * i + 1));
    const int64_t t = (n8 << 3) + (int64_t)blockIdx.x * kThreads + threadIdx.x;      // ragged tail (< 8 elements)
    if (t < n) {
        float x = __fdiv_rn(w[t], beta);
        if (v_out != nullptr) v_out[t] = x;
        if (PEER)
            for (int p = 0; p < vt.count; ++p) vt.v[p][shard_lo + t] = x;
        if (BF16) {
            __nv_bfloat16 h = __float2bfloat16_rn(x);
            row_bf16[t] = *reinterpret_cast<uint16_t*>(&h);
        }
    }
    if (PEER && pv.world > 1 && vt.count > 0) {
        __shared__ unsigned s_ticket_n;
        __threadfence_system();                            // this thread's peer stores are ordered before the ticket
        __syncthreads();
        if (threadIdx.x == 0) s_ticket_n = atomicInc(counter, gridDim.x - 1);
        __syncthreads();
        if (s_ticket_n != gridDim.x - 1) return;
        __threadfence_system();
        peer_push(pv, HLV_CH_V, nullptr, 0);
    }
}

}  // namespace hlv

using namespace hlv;

extern "C" {

int hlv_gather_f32(const void* const* h_src, const int64_t* h_numel, int ntensors, float* dst,
                   int64_t dst_len, float scale, int accumulate, const float* v, double* dot_out,
                   void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return multi_tensor<true>(h_src, h_numel, ntensors, dst, dst_len, scale, accumulate, v, dot_out, ws,
                              ws_bytes, static_cast<cudaStream_t>(stream), "hlv_gather_f32");
}

int hlv_scatter_f32(const float* src, int64_t src_len, void* const* h_dst, const int64_t* h_numel,
                    int ntensors, hlv_stream_t stream) {
    return multi_tensor<false>(const_cast<const void* const*>(h_dst), h_numel, ntensors,
                               const_cast<float*>(src), src_len, 1.0f, 0, nullptr, nullptr, nullptr, 0,
                               static_cast<cudaStream_t>(stream), "hlv_scatter_f32");
}

int hlv_dot_f32(const float* a, const float* b, int64_t n, double* out, void* ws_raw, size_t ws_bytes,
                hlv_stream_t stream) {
    HLV_REQUIRE(a && b && out && n >= 0, HLV_ERR_ARG, "hlv_dot_f32: bad argument");
    HLV_REQUIRE(aligned16(a) && aligned16(b), HLV_ERR_ALIGN, "hlv_dot_f32: operands must be 16-byte aligned");
    Workspace ws;
    HLV_REQUIRE(carve_workspace(ws_raw, ws_bytes, 1, &ws), HLV_ERR_WORKSPACE, "hlv_dot_f32: workspace too small");
    HLV_REQUIRE(sm_count() > 0, HLV_ERR_NO_DEVICE, "hlv_dot_f32: no CUDA device");
    const int grid = persistent_grid((n / 4 + kThreads * kVecPerThread - 1) / (kThreads * kVecPerThread) + 1,
                                     resident_ctas(dot_kernel));
    dot_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, ws.partials, ws.counters, out);
    HLV_LAUNCH_CHECK("hlv_dot_f32 launch");
    return HLV_OK;
}

static int lanczos_update_impl(const char* name, const hlv_peer_ctx* h_ctx, float* w, const float* vj, const float* vjm1, double* alpha,
                               const double* beta, int64_t n, double* norm2_out, void* ws_raw, size_t ws_bytes, hlv_stream_t stream) {
    HLV_REQUIRE(w && vj && alpha && norm2_out && n >= 0, HLV_ERR_ARG, "%s: bad argument", name);
    HLV_REQUIRE((vjm1 == nullptr) == (beta == nullptr), HLV_ERR_ARG, "%s: vjm1 and beta must both be set or both NULL", name);
    HLV_REQUIRE(aligned16(w) && aligned16(vj) && aligned16(vjm1), HLV_ERR_ALIGN, "%s: vectors must be 16-byte aligned", name);
    int rc = check_peer_ctx(h_ctx, name);
    if (rc != HLV_OK) return rc;
    Workspace ws;
    HLV_REQUIRE(carve_workspace(ws_raw, ws_bytes, 1, &ws), HLV_ERR_WORKSPACE, "%s: workspace too small", name);
    HLV_REQUIRE(sm_count() > 0, HLV_ERR_NO_DEVICE, "%s: no CUDA device", name);
    const int grid = persistent_grid((n / 4 + kThreads * kVecPerThread - 1) / (kThreads * kVecPerThread) + 1,
                                     vjm1 ? resident_ctas(lanczos_update_kernel<true>) : resident_ctas(lanczos_update_kernel<false>));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const PeerView pv = make_peer_view(h_ctx);
    if (vjm1)
        lanczos_update_kernel<true><<<grid, kThreads, 0, s>>>(w, vj, vjm1, alpha, beta, n, ws.partials, ws.counters, norm2_out, pv);
    else
        lanczos_update_kernel<false><<<grid, kThreads, 0, s>>>(w, vj, nullptr, alpha, nullptr, n, ws.partials, ws.counters, norm2_out, pv);
    HLV_LAUNCH_CHECK(name);
    return HLV_OK;
}

int hlv_lanczos_update_f32(float* w, const float* vj, const float* vjm1, const double* alpha,
                           const double* beta, int64_t n, double* norm2_out, void* ws_raw, size_t ws_bytes,
                           hlv_stream_t stream) {
    return lanczos_update_impl("hlv_lanczos_update_f32", nullptr, w, vj, vjm1, const_cast<double*>(alpha), beta, n, norm2_out,
                               ws_raw, ws_bytes, stream);
}
int hlv_x_lanczos_update_f32(const hlv_peer_ctx* h_ctx, float* w, const float* vj, const float* vjm1, double* alpha,
                             const double* beta, int64_t n, double* norm2_out, void* ws_raw, size_t ws_bytes,
                             hlv_stream_t stream) {
    return lanczos_update_impl("hlv_x_lanczos_update_f32", h_ctx, w, vj, vjm1, alpha, beta, n, norm2_out, ws_raw, ws_bytes, stream);
}

static int normalize_store_impl(const char* name, const hlv_peer_ctx* h_ctx, const float* w, double* norm2, int64_t n, double* beta_out,
                                float* v_out, uint16_t* row_bf16, float* const* h_v_full, float* v_multicast, int64_t shard_lo,
                                double breakdown_tol, int* breakdown_iter, int iter, void* ws_raw, size_t ws_bytes, hlv_stream_t stream) {
    HLV_REQUIRE(w && norm2 && beta_out && n >= 0, HLV_ERR_ARG, "%s: bad argument", name);
    HLV_REQUIRE(w != v_out, HLV_ERR_ARG, "%s: v_out must not alias w", name);
    int rc = check_peer_ctx(h_ctx, name);
    if (rc != HLV_OK) return rc;
    const PeerView pv = make_peer_view(h_ctx);
    VecTable vt{};
    if (pv.world > 1 && h_v_full != nullptr && (v_out || row_bf16)) {
        HLV_REQUIRE(shard_lo >= 0 && (shard_lo & 3) == 0, HLV_ERR_ALIGN, "%s: shard_lo must be a non-negative multiple of 4", name);
        for (int p = 0; p < pv.world; ++p) {
            HLV_REQUIRE(h_v_full[p] != nullptr && aligned16(h_v_full[p]), HLV_ERR_ALIGN, "%s: full vector of rank %d NULL or unaligned", name, p);
            vt.v[p] = h_v_full[p];
        }
        vt.count = pv.world;
        if (v_multicast != nullptr && aligned16(v_multicast) && (n & 7) == 0) vt.mc = v_multicast;   // ragged shards keep the peer stores
    }
    if (!v_out && !row_bf16) n = 0;                     // beta only (last iteration: residual norm)
    HLV_REQUIRE(aligned16(w) && aligned16(v_out) && aligned16(row_bf16), HLV_ERR_ALIGN, "%s: vectors must be 16-byte aligned", name);
    HLV_REQUIRE(sm_count() > 0, HLV_ERR_NO_DEVICE, "%s: no CUDA device", name);
    Workspace ws{};
    const bool peer = pv.world > 1;
    if (peer) HLV_REQUIRE(carve_workspace(ws_raw, ws_bytes, 1, &ws), HLV_ERR_WORKSPACE, "%s: workspace too small", name);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int64_t items = (n / 8 + kThreads * 2 - 1) / (kThreads * 2) + 1;
    // Peer stores want exactly ONE CTA per SM: measured on 2 x B200 (248 MB to the peer, scripts/dev/peer_store_bench.py) 0.39 ms
    // = 635 GB/s with 148 CTAs against 0.85 ms with the 592 CTAs of the occupancy-sized grid (and 0.53 ms with 200, 0.57 with 96);
    // NCCL's all-gather of the same bytes takes 0.58 ms.  Peer LOADS (the reduce-scatter kernel) want the full grid.
    const int store_grid_cap = (peer && vt.count > 0) ? sm_count() : INT32_MAX;
#define HLV_NORM_LAUNCH(B, P)                                                                                          \
    do {                                                                                                               \
        int grid = persistent_grid(items, resident_ctas(normalize_store_kernel<B, P>));                                \
        if (grid > store_grid_cap) grid = store_grid_cap;                                                              \
        normalize_store_kernel<B, P><<<grid, kThreads, 0, s>>>(w, norm2, n, beta_out, v_out, row_bf16, breakdown_tol,  \
                                                               breakdown_iter, iter, pv, vt, shard_lo, ws.counters);   \
    } while (0)
    if (row_bf16) { if (peer) HLV_NORM_LAUNCH(true, true); else HLV_NORM_LAUNCH(true, false); }
    else          { if (peer) HLV_NORM_LAUNCH(false, true); else HLV_NORM_LAUNCH(false, false); }
#undef HLV_NORM_LAUNCH
    HLV_LAUNCH_CHECK(name);
    return HLV_OK;
}

int hlv_normalize_store_f32(const float* w, const double* norm2, int64_t n, double* beta_out, float* v_out,
                            uint16_t* row_bf16, double breakdown_tol, int* breakdown_iter, int iter,
                            hlv_stream_t stream) {
    return normalize_store_impl("hlv_normalize_store_f32", nullptr, w, const_cast<double*>(norm2), n, beta_out, v_out, row_bf16, nullptr, nullptr, 0,
                                breakdown_tol, breakdown_iter, iter, nullptr, 0, stream);
}
int hlv_x_normalize_store_f32(const hlv_peer_ctx* h_ctx, const float* w, double* norm2, int64_t n, double* beta_out, float* v_out,
                              uint16_t* row_bf16, float* const* h_v_full, float* v_multicast, int64_t shard_lo, double breakdown_tol,
                              int* breakdown_iter, int iter, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return normalize_store_impl("hlv_x_normalize_store_f32", h_ctx, w, norm2, n, beta_out, v_out, row_bf16, h_v_full, v_multicast, shard_lo,
                                breakdown_tol, breakdown_iter, iter, ws, ws_bytes, stream);
}

}  // extern "C"
