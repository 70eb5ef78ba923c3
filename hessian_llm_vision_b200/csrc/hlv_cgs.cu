// libhlv.so -- kernel (c): classical Gram-Schmidt against a row-major basis resident in HBM,
// as a bandwidth-bound tall-skinny GEMV pair
//     project:  c = V w        (rows dot products, ONE streaming pass over V)
//     update :  w += sign * V^T c  (+ sum w^2 in the same pass, second pass over V)
// plus the two consumers that are the same memory pattern: the low-rank gradient adjustment
// (drop-in for the reference's vector_adjust.cu) and Ritz-vector materialisation.
//
// Work decomposition ("column owner"): a CTA owns column tiles of kTile consecutive elements;
// per tile it keeps its slice of w in registers and streams the `rows` basis rows of that tile
// past it, 8 fp32 / 16 bf16 rows (= 16 independent 128-bit loads per thread) at a time.
// V is read exactly once per pass, w once (project) or once + one write (update).
// Arithmetic intensity is 0.5 flop/byte (fp32 basis) -- HBM roofline, no tensor cores.
#include "hlv_peer.cuh"

namespace hlv {

constexpr int kEpt = 8;                         // elements of w per thread per tile
constexpr int kTile = kThreads * kEpt;          // 2048 columns per tile (8 KB of an fp32 row)
// rows whose loads are in flight together: 16 x 128-bit loads per thread for either storage type
template <typename BT> __host__ __device__ constexpr int row_batch() { return sizeof(BT) == 2 ? 16 : 8; }

// ---- per-thread slice of one basis row: 8 consecutive-by-4 elements ---------------------
// Element layout inside a tile keeps every 128-bit access fully coalesced:
//   fp32 : vec k (k=0,1) of thread t covers columns  k*1024 + 4t .. +3
//   bf16 : one 128-bit load of thread t covers columns 8t .. 8t+7
template <typename BT> struct RowSlice;

template <> struct RowSlice<float> {
    float4 a, b;
    static __device__ __forceinline__ int col(int tid, int e) { return (e >> 2) * (kThreads * 4) + tid * 4 + (e & 3); }
    __device__ __forceinline__ void load(const float* row_tile, int tid) {
        a = ldg_stream(reinterpret_cast<const float4*>(row_tile) + tid);
        b = ldg_stream(reinterpret_cast<const float4*>(row_tile) + kThreads + tid);
    }
    __device__ __forceinline__ void load_guarded(const float* row_tile, int tid, int64_t valid) {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { int c = col(tid, e); t[e] = c < valid ? row_tile[c] : 0.0f; }
        a = make_float4(t[0], t[1], t[2], t[3]); b = make_float4(t[4], t[5], t[6], t[7]);
    }
    __device__ __forceinline__ void unpack(float (&x)[8]) const {
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
};

template <> struct RowSlice<uint16_t> {
    uint4 a;
    static __device__ __forceinline__ int col(int tid, int e) { return tid * 8 + e; }
    __device__ __forceinline__ void load(const uint16_t* row_tile, int tid) {
        a = ldg_stream(reinterpret_cast<const uint4*>(row_tile) + tid);
    }
    __device__ __forceinline__ void load_guarded(const uint16_t* row_tile, int tid, int64_t valid) {
        uint32_t t[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            int c = tid * 8 + 2 * p;
            uint32_t lo = c < valid ? row_tile[c] : 0u, hi = (c + 1) < valid ? row_tile[c + 1] : 0u;
            t[p] = lo | (hi << 16);
        }
        a = make_uint4(t[0], t[1], t[2], t[3]);
    }
    __device__ __forceinline__ void unpack(float (&x)[8]) const {
        x[0] = bf16_lo(a.x); x[1] = bf16_hi(a.x); x[2] = bf16_lo(a.y); x[3] = bf16_hi(a.y);
        x[4] = bf16_lo(a.z); x[5] = bf16_hi(a.z); x[6] = bf16_lo(a.w); x[7] = bf16_hi(a.w);
    }
};

// w slice of a thread, in the SAME column mapping as RowSlice<BT>.
template <typename BT>
__device__ __forceinline__ void load_w(const float* w_tile, int tid, int64_t valid, float (&x)[8]) {
    if (valid >= kTile) {
        if (sizeof(BT) == 4) {
            float4 a = *(reinterpret_cast<const float4*>(w_tile) + tid);
            float4 b = *(reinterpret_cast<const float4*>(w_tile) + kThreads + tid);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
            float4 a = *(reinterpret_cast<const float4*>(w_tile) + 2 * tid);
            float4 b = *(reinterpret_cast<const float4*>(w_tile) + 2 * tid + 1);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) { int c = RowSlice<BT>::col(tid, e); x[e] = c < valid ? w_tile[c] : 0.0f; }
    }
}
template <typename BT>
__device__ __forceinline__ void store_w(float* w_tile, int tid, int64_t valid, const float (&x)[8]) {
    if (valid >= kTile) {
        float4 a = make_float4(x[0], x[1], x[2], x[3]), b = make_float4(x[4], x[5], x[6], x[7]);
        if (sizeof(BT) == 4) {
            *(reinterpret_cast<float4*>(w_tile) + tid) = a;
            *(reinterpret_cast<float4*>(w_tile) + kThreads + tid) = b;
        } else {
            *(reinterpret_cast<float4*>(w_tile) + 2 * tid) = a;
            *(reinterpret_cast<float4*>(w_tile) + 2 * tid + 1) = b;
        }
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) { int c = RowSlice<BT>::col(tid, e); if (c < valid) w_tile[c] = x[e]; }
    }
}

// =============================================================================
// project: c[i] = <V_i, w>
// UPD: the three-term update  w -= alpha*vj + beta*vjm1  (lanczostrain_hand.py:202, torch's rounding sequence:
// two products, one sum, one subtraction, no FMA contraction) is applied to the CTA's register slice of w first and
// written back, then the rows are streamed past the UPDATED slice -- one launch and 12n bytes less than the separate
// update kernel (v_j and v_{j-1} are rows of the basis that this pass reads anyway; they come back from L2).
// With a peer view: alpha = rank-ordered total of HLV_CH_ALPHA, and the row totals of this rank are pushed on HLV_CH_C1.
// =============================================================================
template <typename BT, bool UPD>
__global__ void __launch_bounds__(kThreads, 2)
cgs_project_kernel(const BT* __restrict__ V, int64_t ldv, int rows, float* __restrict__ w, int64_t n,
                   const float* __restrict__ vj, const float* __restrict__ vjm1, double* alpha_p, const double* __restrict__ beta_p,
                   double* partials, unsigned* counter, double* c_out, const __grid_constant__ PeerView pv, int push_channel) {
    extern __shared__ double s_acc[];                   // [kWarps][rows_pad]: per-warp running sums
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kRowBatch = row_batch<BT>();
    const int rows_pad = (rows + kRowBatch - 1) / kRowBatch * kRowBatch;
    float alpha = 0.0f, beta = 0.0f;
    if (UPD) {
        double a;
        if (pv.world > 1) {
            a = peer_pull_scalar(pv, HLV_CH_ALPHA);
            if (blockIdx.x == 0 && tid == 0) alpha_p[0] = a;      // the total, for the host's T
        } else {
            a = alpha_p[0];
        }
        alpha = (float)a;
        beta = vjm1 != nullptr ? (float)beta_p[0] : 0.0f;
    }
    for (int i = tid; i < kWarps * rows_pad; i += kThreads) s_acc[i] = 0.0;
    __syncthreads();
    double* my_acc = s_acc + warp * rows_pad;
    const int my_row = warp_sum8_row(lane);
    const int64_t ntiles = (n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t x0 = tile * kTile;
        const int64_t valid = n - x0;                   // >= kTile for every tile but a ragged last one
        float wv[8];
        load_w<BT>(w + x0, tid, valid, wv);
        if (UPD) {
            float a[8], b[8];
            load_w<BT>(vj + x0, tid, valid, a);
            if (vjm1 != nullptr) load_w<BT>(vjm1 + x0, tid, valid, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float t = __fmul_rn(alpha, a[e]);
                if (vjm1 != nullptr) t = __fadd_rn(t, __fmul_rn(beta, b[e]));
                wv[e] = __fsub_rn(wv[e], t);
            }
            store_w<BT>(w + x0, tid, valid, wv);
        }
        const BT* col0 = V + x0;
        for (int r0 = 0; r0 < rows; r0 += kRowBatch) {
            RowSlice<BT> s[kRowBatch];
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i) {
                if (r0 + i < rows) {
                    const BT* row_tile = col0 + (int64_t)(r0 + i) * ldv;
                    if (valid >= kTile) s[i].load(row_tile, tid); else s[i].load_guarded(row_tile, tid, valid);
                }
            }
            float p[kRowBatch];
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i) {
                p[i] = 0.0f;
                if (r0 + i < rows) {
                    float x[8];
                    s[i].unpack(x);
#pragma unroll
                    for (int e = 0; e < 8; ++e) p[i] = fmaf(x[e], wv[e], p[i]);
                }
            }
#pragma unroll
            for (int h = 0; h < kRowBatch / 8; ++h) {
                const float tot = warp_sum8(*reinterpret_cast<float(*)[8]>(p + 8 * h), lane);
                if ((lane & 3) == 0) my_acc[r0 + 8 * h + my_row] += (double)tot;   // rows_pad covers the index
            }
        }
    }
    __syncthreads();
    for (int r = tid; r < rows; r += kThreads) {
        double t = 0.0;
#pragma unroll
        for (int wi = 0; wi < kWarps; ++wi) t += s_acc[wi * rows_pad + r];
        partials[(size_t)r * kMaxCtas + blockIdx.x] = t;
    }
    finalize_rows_push(partials, counter, rows, c_out, pv, push_channel);
}

// =============================================================================
// update: w += sign * sum_i c[i] V_i ; norm2 = sum w^2
// =============================================================================
template <typename BT>
__global__ void __launch_bounds__(kThreads, 2)
cgs_update_kernel(const BT* __restrict__ V, int64_t ldv, int rows, double* c, float sign,
                  float* __restrict__ w, int64_t n, double* partials, unsigned* counter, double* norm2_out,
                  const int* __restrict__ run_flag, const __grid_constant__ PeerView pv) {
    extern __shared__ float s_c[];                      // sign * (float)c[i]
    __shared__ double s_warp[kWarps];
    const int tid = threadIdx.x;
    // device-side predication: a pass that an earlier kernel found unnecessary costs one launch and no traffic;
    // w and norm2_out are left untouched and the ticket counter is not taken (every CTA leaves here)
    if (run_flag != nullptr && *run_flag == 0) return;
    constexpr int kRowBatch = row_batch<BT>();
    if (pv.world > 1) {                                 // c = rank-ordered totals of the second projection
        const bool keep = blockIdx.x == 0;
        peer_pull(pv, HLV_CH_C2, rows, [&](int i, double t) { s_c[i] = sign * (float)t; if (keep) c[i] = t; });
    } else {
        for (int i = tid; i < rows; i += kThreads) s_c[i] = sign * (float)c[i];
        __syncthreads();
    }
    float nrm = 0.0f;
    const int64_t ntiles = (n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t x0 = tile * kTile;
        const int64_t valid = n - x0;
        float acc[8];
        load_w<BT>(w + x0, tid, valid, acc);
        const BT* col0 = V + x0;
        for (int r0 = 0; r0 < rows; r0 += kRowBatch) {
            RowSlice<BT> s[kRowBatch];
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i) {
                if (r0 + i < rows) {
                    const BT* row_tile = col0 + (int64_t)(r0 + i) * ldv;
                    if (valid >= kTile) s[i].load(row_tile, tid); else s[i].load_guarded(row_tile, tid, valid);
                }
            }
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i) {
                if (r0 + i < rows) {
                    const float ci = s_c[r0 + i];
                    float x[8];
                    s[i].unpack(x);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] = fmaf(ci, x[e], acc[e]);
                }
            }
        }
        store_w<BT>(w + x0, tid, valid, acc);
#pragma unroll
        for (int e = 0; e < 8; ++e) nrm = fmaf(acc[e], acc[e], nrm);   // out-of-range slots hold 0
    }
    if (norm2_out != nullptr) {
        double t = block_sum((double)nrm, s_warp);
        if (tid == 0) partials[blockIdx.x] = t;
        finalize_rows_push(partials, counter, 1, norm2_out, pv, HLV_CH_NORM);
    }
}

// flag = 1 if another Gram-Schmidt pass is needed: some |c[i]| > tol * |w|  (gpytorch's "while any q_i . r > tol",
// SURVEY Appendix B, with r normalised); NaNs ask for the pass.  One CTA; rows <= HLV_MAX_ROWS.
__global__ void cgs_needs_pass_kernel(const double* __restrict__ c, int rows, const double* __restrict__ norm2, double tol,
                                      int* flag) {
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    __syncthreads();
    const double lim = tol * sqrt(norm2[0]);
    int any = 0;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) any |= !(fabs(c[i]) <= lim);
    if (any) atomicOr(&s_any, 1);
    __syncthreads();
    if (threadIdx.x == 0) flag[0] = s_any;
}

// coef[i] = (1/eig[i] - 1/(eig[i]+delta)) * dots[i]   (vector_adjust.cu:11, fp32 like the reference)
__global__ void adjust_coef_kernel(const double* dots, const float* eigvals, float delta, int k, double* coef) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) {
        float lam = eigvals[i];
        float s = 1.0f / lam - 1.0f / (lam + delta);
        coef[i] = (double)(s * (float)dots[i]);
    }
}

// =============================================================================
// Ritz vectors: out[v] = sum_i Y[i, v0+v] Q_i for kNv outputs per pass over Q
// =============================================================================
constexpr int kNv = 8;
constexpr int kRitzBatch = 4;

template <typename BT>
__global__ void __launch_bounds__(kThreads, 2)
ritz_vectors_kernel(const BT* __restrict__ Q, int64_t ldq, int m, const float* __restrict__ Y, int ldy,
                    int v0, int nv, float* __restrict__ out, int64_t ldo, int64_t n) {
    extern __shared__ float s_y[];                      // [m][kNv]
    const int tid = threadIdx.x;
    for (int i = tid; i < m * kNv; i += kThreads) {
        int r = i / kNv, v = i % kNv;
        s_y[i] = v < nv ? Y[(int64_t)r * ldy + v0 + v] : 0.0f;
    }
    __syncthreads();
    const int64_t ntiles = (n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t x0 = tile * kTile;
        const int64_t valid = n - x0;
        float acc[kNv][8];
#pragma unroll
        for (int v = 0; v < kNv; ++v)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[v][e] = 0.0f;
        const BT* col0 = Q + x0;
        for (int r0 = 0; r0 < m; r0 += kRitzBatch) {
            RowSlice<BT> s[kRitzBatch];
#pragma unroll
            for (int i = 0; i < kRitzBatch; ++i) {
                if (r0 + i < m) {
                    const BT* row_tile = col0 + (int64_t)(r0 + i) * ldq;
                    if (valid >= kTile) s[i].load(row_tile, tid); else s[i].load_guarded(row_tile, tid, valid);
                }
            }
#pragma unroll
            for (int i = 0; i < kRitzBatch; ++i) {
                if (r0 + i < m) {
                    float x[8];
                    s[i].unpack(x);
                    const float4 y0 = *reinterpret_cast<const float4*>(s_y + (r0 + i) * kNv);
                    const float4 y1 = *reinterpret_cast<const float4*>(s_y + (r0 + i) * kNv + 4);
                    const float yy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
                    for (int v = 0; v < kNv; ++v)
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[v][e] = fmaf(yy[v], x[e], acc[v][e]);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < kNv; ++v)
            if (v < nv) store_w<BT>(out + (int64_t)(v0 + v) * ldo + x0, tid, valid, acc[v]);
    }
}

// ---- host launchers ---------------------------------------------------------------------
template <typename BT>
static int check_basis(const char* name, const BT* V, int64_t ldv, int rows, const float* w, int64_t n) {
    HLV_REQUIRE(V && w && n >= 0 && rows >= 1 && rows <= HLV_MAX_ROWS, HLV_ERR_ARG,
                "%s: bad argument (rows=%d, n=%lld)", name, rows, (long long)n);
    HLV_REQUIRE(ldv >= n, HLV_ERR_ARG, "%s: ldv=%lld < n=%lld", name, (long long)ldv, (long long)n);
    HLV_REQUIRE(aligned16(V) && aligned16(w) && ((ldv * (int64_t)sizeof(BT)) & 15) == 0, HLV_ERR_ALIGN,
                "%s: V, w must be 16-byte aligned and ldv*sizeof(elem) a multiple of 16", name);
    HLV_REQUIRE(sm_count() > 0, HLV_ERR_NO_DEVICE, "%s: no CUDA device", name);
    return HLV_OK;
}

template <typename BT, bool UPD>
static int project(const char* name, const hlv_peer_ctx* h_ctx, const BT* V, int64_t ldv, int rows, float* w, int64_t n,
                   const float* vj, const float* vjm1, double* alpha, const double* beta,
                   double* c_out, void* ws_raw, size_t ws_bytes, cudaStream_t stream) {
    int rc = check_basis(name, V, ldv, rows, w, n);
    if (rc != HLV_OK) return rc;
    if ((rc = check_peer_ctx(h_ctx, name)) != HLV_OK) return rc;
    HLV_REQUIRE(c_out != nullptr, HLV_ERR_ARG, "%s: c_out is NULL", name);
    if (UPD) {
        HLV_REQUIRE(vj && alpha && ((vjm1 == nullptr) == (beta == nullptr)), HLV_ERR_ARG,
                    "%s: vj and alpha are required; vjm1 and beta go together", name);
        HLV_REQUIRE(aligned16(vj) && aligned16(vjm1), HLV_ERR_ALIGN, "%s: vj, vjm1 must be 16-byte aligned", name);
    }
    Workspace ws;
    HLV_REQUIRE(carve_workspace(ws_raw, ws_bytes, rows, &ws), HLV_ERR_WORKSPACE,
                "%s: workspace too small for %d rows (need %zu bytes)", name, rows, workspace_bytes(rows));
    const int rows_pad = (rows + row_batch<BT>() - 1) / row_batch<BT>() * row_batch<BT>();
    const size_t smem = (size_t)kWarps * rows_pad * sizeof(double);
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void*>(cgs_project_kernel<BT, UPD>),
                                        smem > 48 * 1024 ? (size_t)kWarps * (HLV_MAX_ROWS + 16) * sizeof(double) : 0);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(project)");
    const int grid = persistent_grid((n + kTile - 1) / kTile, resident_ctas(cgs_project_kernel<BT, UPD>, smem));
    const PeerView pv = make_peer_view(h_ctx);
    cgs_project_kernel<BT, UPD><<<grid, kThreads, smem, stream>>>(V, ldv, rows, w, n, vj, vjm1, alpha, beta, ws.partials, ws.counters,
                                                                   c_out, pv, pv.world > 1 ? HLV_CH_C1 : -1);
    HLV_LAUNCH_CHECK(name);
    return HLV_OK;
}

template <typename BT>
static int update(const char* name, const hlv_peer_ctx* h_ctx, const BT* V, int64_t ldv, int rows, double* c, float sign, float* w,
                  int64_t n, double* norm2_out, void* ws_raw, size_t ws_bytes, cudaStream_t stream,
                  const int* run_flag = nullptr) {
    int rc = check_basis(name, V, ldv, rows, w, n);
    if (rc != HLV_OK) return rc;
    if ((rc = check_peer_ctx(h_ctx, name)) != HLV_OK) return rc;
    HLV_REQUIRE(c != nullptr, HLV_ERR_ARG, "%s: c is NULL", name);
    const PeerView pv = make_peer_view(h_ctx);
    HLV_REQUIRE(pv.world == 1 || (norm2_out != nullptr && run_flag == nullptr), HLV_ERR_ARG,
                "%s: with a peer context the pass always runs and always reduces |w|^2", name);
    Workspace ws{};
    if (norm2_out)
        HLV_REQUIRE(carve_workspace(ws_raw, ws_bytes, 1, &ws), HLV_ERR_WORKSPACE, "%s: workspace too small", name);
    const int grid = persistent_grid((n + kTile - 1) / kTile, resident_ctas(cgs_update_kernel<BT>, rows * sizeof(float)));
    cgs_update_kernel<BT><<<grid, kThreads, rows * sizeof(float), stream>>>(V, ldv, rows, c, sign, w, n, ws.partials,
                                                                           ws.counters, norm2_out, run_flag, pv);
    HLV_LAUNCH_CHECK(name);
    return HLV_OK;
}

// hlv_ritz_tc.cu: the tensor-core pass over the columns [0, n_main)
int ritz_vectors_tc(const float* Q, int64_t ldq, int m, const float* Y, int ldy, int v0, int nvec, float* out, int64_t ldo,
                    int64_t n, cudaStream_t stream, int64_t* n_main_out);

template <typename BT>
static int ritz_vectors(const char* name, const BT* Q, int64_t ldq, int m, const float* Y, int ldy, int nvec,
                        float* out, int64_t ldo, int64_t n, cudaStream_t stream) {
    HLV_REQUIRE(Q && Y && out && n >= 0 && m >= 1 && m <= HLV_MAX_ROWS && nvec >= 0 && ldy >= nvec, HLV_ERR_ARG,
                "%s: bad argument", name);
    HLV_REQUIRE(ldq >= n && ldo >= n, HLV_ERR_ARG, "%s: leading dimension < n", name);
    HLV_REQUIRE(aligned16(Q) && aligned16(out) && ((ldq * (int64_t)sizeof(BT)) & 15) == 0 && ((ldo * 4) & 15) == 0,
                HLV_ERR_ALIGN, "%s: Q, out must be 16-byte aligned with 16-byte row pitch", name);
    HLV_REQUIRE(sm_count() > 0, HLV_ERR_NO_DEVICE, "%s: no CUDA device", name);
    // More than 8 vectors from an fp32 basis: ONE pass over Q on the tensor cores (3xTF32, hlv_ritz_tc.cu), up to 112 vectors
    // per pass; the CUDA-core kernel below (8 vectors per pass over Q) keeps the ragged last < 128 columns, small requests
    // and bf16 rows.
    for (int v0 = 0; v0 < nvec;) {
        int64_t n_main = 0;
        int slice = nvec - v0 < 112 ? nvec - v0 : 112;
        if (sizeof(BT) == 4 && slice > kNv) {
            const int rc = ritz_vectors_tc(reinterpret_cast<const float*>(Q), ldq, m, Y, ldy, v0, slice, out, ldo, n, stream, &n_main);
            if (rc != HLV_OK) return rc;
        }
        if (n_main == 0) slice = slice < kNv ? slice : kNv;                       // CUDA-core pass over everything: 8 at a time
        if (n_main < n) {
            const int64_t nt = n - n_main;
            const int grid = persistent_grid((nt + kTile - 1) / kTile, resident_ctas(ritz_vectors_kernel<BT>, (size_t)m * kNv * sizeof(float)));
            for (int u0 = v0; u0 < v0 + slice; u0 += kNv) {
                const int nv = v0 + slice - u0 < kNv ? v0 + slice - u0 : kNv;
                ritz_vectors_kernel<BT><<<grid, kThreads, (size_t)m * kNv * sizeof(float), stream>>>(Q + n_main, ldq, m, Y, ldy, u0, nv,
                                                                                                    out + n_main, ldo, nt);
                HLV_LAUNCH_CHECK(name);
            }
        }
        v0 += slice;
    }
    return HLV_OK;
}

}  // namespace hlv

using namespace hlv;

extern "C" {

int hlv_cgs_project_f32(const float* V, int64_t ldv, int rows, const float* w, int64_t n, double* c_out,
                        void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return project<float, false>("hlv_cgs_project_f32", nullptr, V, ldv, rows, const_cast<float*>(w), n, nullptr, nullptr, nullptr, nullptr,
                                 c_out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int hlv_cgs_project_bf16(const uint16_t* V, int64_t ldv, int rows, const float* w, int64_t n, double* c_out,
                         void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return project<uint16_t, false>("hlv_cgs_project_bf16", nullptr, V, ldv, rows, const_cast<float*>(w), n, nullptr, nullptr, nullptr, nullptr,
                                    c_out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int hlv_cgs_update_f32(const float* V, int64_t ldv, int rows, const double* c, float sign, float* w, int64_t n,
                       double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update<float>("hlv_cgs_update_f32", nullptr, V, ldv, rows, const_cast<double*>(c), sign, w, n, norm2_out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int hlv_cgs_update_bf16(const uint16_t* V, int64_t ldv, int rows, const double* c, float sign, float* w, int64_t n,
                        double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update<uint16_t>("hlv_cgs_update_bf16", nullptr, V, ldv, rows, const_cast<double*>(c), sign, w, n, norm2_out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int hlv_cgs_update_if_f32(const float* V, int64_t ldv, int rows, const double* c, float sign, float* w, int64_t n,
                          double* norm2_out, const int* run_flag, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update<float>("hlv_cgs_update_if_f32", nullptr, V, ldv, rows, const_cast<double*>(c), sign, w, n, norm2_out, ws, ws_bytes, static_cast<cudaStream_t>(stream), run_flag);
}
int hlv_cgs_update_if_bf16(const uint16_t* V, int64_t ldv, int rows, const double* c, float sign, float* w, int64_t n,
                           double* norm2_out, const int* run_flag, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update<uint16_t>("hlv_cgs_update_if_bf16", nullptr, V, ldv, rows, const_cast<double*>(c), sign, w, n, norm2_out, ws, ws_bytes, static_cast<cudaStream_t>(stream), run_flag);
}
int hlv_cgs_needs_pass(const double* c, int rows, const double* norm2, double tol, int* flag_out, hlv_stream_t stream) {
    HLV_REQUIRE(c && norm2 && flag_out && rows >= 1 && tol >= 0.0, HLV_ERR_ARG, "hlv_cgs_needs_pass: bad argument");
    cgs_needs_pass_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(c, rows, norm2, tol, flag_out);
    HLV_LAUNCH_CHECK("hlv_cgs_needs_pass");
    return HLV_OK;
}

int hlv_x_update_project_f32(const hlv_peer_ctx* h_ctx, const float* V, int64_t ldv, int rows, float* w, int64_t n,
                             const float* vj, const float* vjm1, double* alpha, const double* beta,
                             double* c_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return project<float, true>("hlv_x_update_project_f32", h_ctx, V, ldv, rows, w, n, vj, vjm1, alpha, beta, c_out, ws, ws_bytes,
                                static_cast<cudaStream_t>(stream));
}
int hlv_x_update_project_bf16(const hlv_peer_ctx* h_ctx, const uint16_t* V, int64_t ldv, int rows, float* w, int64_t n,
                              const float* vj, const float* vjm1, double* alpha, const double* beta,
                              double* c_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return project<uint16_t, true>("hlv_x_update_project_bf16", h_ctx, V, ldv, rows, w, n, vj, vjm1, alpha, beta, c_out, ws, ws_bytes,
                                   static_cast<cudaStream_t>(stream));
}
int hlv_x_cgs_update_f32(const hlv_peer_ctx* h_ctx, const float* V, int64_t ldv, int rows, double* c, float* w, int64_t n,
                         double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update<float>("hlv_x_cgs_update_f32", h_ctx, V, ldv, rows, c, -1.0f, w, n, norm2_out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
int hlv_x_cgs_update_bf16(const hlv_peer_ctx* h_ctx, const uint16_t* V, int64_t ldv, int rows, double* c, float* w, int64_t n,
                          double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update<uint16_t>("hlv_x_cgs_update_bf16", h_ctx, V, ldv, rows, c, -1.0f, w, n, norm2_out, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int hlv_vector_adjust_f32(const float* grad_vector, const float* V, const float* eigvals,
                          float* adjusted_grad_vector, int num_eigenvalues, int64_t vec_len, float delta,
                          int64_t ldv, double* coef_scratch, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    HLV_REQUIRE(eigvals && coef_scratch && adjusted_grad_vector, HLV_ERR_ARG, "hlv_vector_adjust_f32: bad argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int rc = project<float, false>("hlv_vector_adjust_f32/project", nullptr, V, ldv, num_eigenvalues, const_cast<float*>(grad_vector), vec_len,
                                   nullptr, nullptr, nullptr, nullptr, coef_scratch, ws, ws_bytes, s);
    if (rc != HLV_OK) return rc;
    adjust_coef_kernel<<<(num_eigenvalues + 127) / 128, 128, 0, s>>>(coef_scratch, eigvals, delta, num_eigenvalues,
                                                                     coef_scratch);
    HLV_LAUNCH_CHECK("hlv_vector_adjust_f32/coef");
    return update<float>("hlv_vector_adjust_f32/update", nullptr, V, ldv, num_eigenvalues, coef_scratch, 1.0f,
                         adjusted_grad_vector, vec_len, nullptr, ws, ws_bytes, s);
}

int hlv_ritz_vectors_f32(const float* Q, int64_t ldq, int m, const float* Y, int ldy, int nvec, float* out,
                         int64_t ldo, int64_t n, hlv_stream_t stream) {
    return ritz_vectors<float>("hlv_ritz_vectors_f32", Q, ldq, m, Y, ldy, nvec, out, ldo, n, static_cast<cudaStream_t>(stream));
}
int hlv_ritz_vectors_bf16(const uint16_t* Q, int64_t ldq, int m, const float* Y, int ldy, int nvec, float* out,
                          int64_t ldo, int64_t n, hlv_stream_t stream) {
    return ritz_vectors<uint16_t>("hlv_ritz_vectors_bf16", Q, ldq, m, Y, ldy, nvec, out, ldo, n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
