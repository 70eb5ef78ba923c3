/*
 * hlv.h -- C ABI of libhlv.so: the B200 (sm_100a) kernels of the Lanczos /
 * stochastic-Lanczos-quadrature recurrence hot path.
 *
 * Boundary being replaced.  The reference's only native code is the bare kernel
 *   extern "C" __global__ void vector_adjust(const float* grad_vector, const float* V,
 *        const float* eigvals, float* adjusted_grad_vector, int num_eigenvalues,
 *        int vec_len, float delta)                       (vector_adjust.cu:2)
 * built with `nvcc --shared -o vector_adjust.so --compiler-options '-fPIC'`
 * (shared_kernel:1) and launched from Python through pycuda with raw device
 * pointers of caller-owned, row-major contiguous torch CUDA tensors
 * (gpt_hessian_cuda.py:25-54).  Everything else on the path is stock torch ops
 * (torch.cat / dot / norm / axpy, and the reorth loop inside gpytorch).  This
 * header keeps that contract -- raw device pointers, caller-owned buffers,
 * row-major fp32, nothing allocated by the library -- and adds what the
 * reference lacks: HOST entry points (not bare kernels), a stream argument, and
 * an error convention.
 *
 * Conventions (every function):
 *   - returns HLV_OK (0) or a negative HLV_ERR_* code; never throws, never
 *     synchronises the device, never allocates device memory;
 *   - all pointers are DEVICE pointers unless the parameter name starts with
 *     `h_` (host arrays that are consumed before the call returns);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     work is stream-ordered;
 *   - scalar results (dot products, squared norms, Gram-Schmidt coefficients)
 *     are written to caller-provided DEVICE doubles: they are fp32 per-thread
 *     FMA chains combined in a fixed order with an fp64 final stage, so a result
 *     depends only on (n, launch geometry), never on block scheduling;
 *   - `ws` is a scratch buffer of at least hlv_workspace_bytes(max_rows) bytes
 *     that must be zeroed ONCE (hlv_workspace_init) and must not be shared by
 *     calls running concurrently on different streams;
 *   - vectors of length n: base pointers 16-byte aligned; basis rows are
 *     row-major with leading dimension `ldv` elements, ldv*sizeof(elem) % 16 == 0.
 *     n itself is arbitrary (ragged tails are handled in-kernel).
 */
#ifndef HLV_H_
#define HLV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HLV_VERSION 200            /* 0.2.0 */

#define HLV_OK              0
#define HLV_ERR_ARG        -1      /* null pointer / negative size / rows out of range */
#define HLV_ERR_ALIGN      -2      /* pointer or leading dimension not 16-byte aligned */
#define HLV_ERR_WORKSPACE  -3      /* ws too small */
#define HLV_ERR_CUDA       -4      /* CUDA runtime error; see hlv_last_error_string() */
#define HLV_ERR_NO_DEVICE  -5      /* no CUDA device / not sm_100 */

#define HLV_MAX_ROWS      1024     /* max basis rows per project/update call */
#define HLV_MAX_TENSORS   1024     /* max tensors per gather/scatter launch (longer lists are chunked) */

#define HLV_MAX_PEERS       16     /* ranks of one NVLink domain that can exchange through peer memory */
#define HLV_PEER_CHANNELS    8

typedef void* hlv_stream_t;

/* ---- library ------------------------------------------------------------ */
int         hlv_version(void);
const char* hlv_last_error_string(void);                 /* thread-local, never NULL */
/* sm count / compute capability of the current device (cached). */
int         hlv_device_info(int* sm_count, int* cc_major, int* cc_minor);
size_t      hlv_workspace_bytes(int max_rows);
int         hlv_workspace_init(void* ws, size_t ws_bytes, hlv_stream_t stream);

/* ---- (a) gather / scatter between per-tensor buffers and the flat vector -- */
/* Replaces torch.cat([p.grad.view(-1) ...]) (gpt2_hessian_cpu.py:109,200) and the
 * slice/split of :79-82,:231-233.  dst[off_t + i] = src_t[i] in list order.
 *   scale/accumulate: dst = (accumulate ? dst : 0) + scale*src.  With scale==1 and
 *   accumulate==0 the copy is bit-exact (no arithmetic is performed).
 *   v, dot_out (both NULL or both set): also writes dot_out[0] = sum_i dst_new[i]*v[i]
 *   (the Lanczos alpha, lanczostrain_hand.py:200) in the same pass.
 *   h_src[t] are device pointers to contiguous fp32 tensors of h_numel[t] elements. */
int hlv_gather_f32(const void* const* h_src, const int64_t* h_numel, int ntensors,
                   float* dst, int64_t dst_len, float scale, int accumulate,
                   const float* v, double* dot_out,
                   void* ws, size_t ws_bytes, hlv_stream_t stream);
/* dst_t[i] = src[off_t + i]  (bit-exact). */
int hlv_scatter_f32(const float* src, int64_t src_len,
                    void* const* h_dst, const int64_t* h_numel, int ntensors,
                    hlv_stream_t stream);

/* ---- (b) recurrence -------------------------------------------------------- */
/* out[0] = sum a[i]*b[i]   (torch.dot, lanczostrain_hand.py:183,200) */
int hlv_dot_f32(const float* a, const float* b, int64_t n, double* out,
                void* ws, size_t ws_bytes, hlv_stream_t stream);
/* w -= alpha*vj + beta*vjm1 and norm2_out[0] = sum w_new^2, one pass
 * (lanczostrain_hand.py:202 fused with :190).  alpha/beta are device doubles,
 * rounded to fp32 before use; vjm1/beta may be NULL (first iteration, :185). */
int hlv_lanczos_update_f32(float* w, const float* vj, const float* vjm1,
                           const double* alpha, const double* beta, int64_t n,
                           double* norm2_out,
                           void* ws, size_t ws_bytes, hlv_stream_t stream);
/* beta = sqrt(norm2[0]); beta_out[0] = beta; v_out = w / beta (lanczostrain_hand.py:190-194).
 * row_bf16 (optional) additionally receives the bf16-rounded copy (bf16 basis storage);
 * v_out may be NULL (bf16 row only); with both NULL only beta_out is written.  v_out != w.
 * If beta < breakdown_tol and *breakdown_iter < 0, *breakdown_iter = iter. */
int hlv_normalize_store_f32(const float* w, const double* norm2, int64_t n,
                            double* beta_out, float* v_out, uint16_t* row_bf16,
                            double breakdown_tol, int* breakdown_iter, int iter,
                            hlv_stream_t stream);

/* ---- (c) classical Gram-Schmidt against a row-major basis in HBM ----------- */
/* c_out[i] = sum_x V[i*ldv + x] * w[x],  i < rows      (one streaming pass over V) */
int hlv_cgs_project_f32 (const float*    V, int64_t ldv, int rows, const float* w, int64_t n,
                         double* c_out, void* ws, size_t ws_bytes, hlv_stream_t stream);
int hlv_cgs_project_bf16(const uint16_t* V, int64_t ldv, int rows, const float* w, int64_t n,
                         double* c_out, void* ws, size_t ws_bytes, hlv_stream_t stream);
/* w[x] += sign * sum_i c[i] * V[i*ldv + x]; norm2_out[0] = sum w_new^2 (may be NULL).
 * c are device doubles, rounded to fp32 before use.  sign = -1 for reorthogonalisation. */
int hlv_cgs_update_f32 (const float*    V, int64_t ldv, int rows, const double* c, float sign,
                        float* w, int64_t n, double* norm2_out,
                        void* ws, size_t ws_bytes, hlv_stream_t stream);
int hlv_cgs_update_bf16(const uint16_t* V, int64_t ldv, int rows, const double* c, float sign,
                        float* w, int64_t n, double* norm2_out,
                        void* ws, size_t ws_bytes, hlv_stream_t stream);

/* Conditional re-orthogonalisation without a host round trip (gpytorch's lanczos_tridiag re-orthogonalises "while any
 * q_i . r > tol", called at gpt2_hessian_cpu.py:207-213).  hlv_cgs_needs_pass sets flag_out[0] = 1
 * when some |c[i]| > tol * sqrt(norm2[0]) (c = the coefficients a projection just measured, norm2 = |w|^2), else 0;
 * hlv_cgs_update_if_* is hlv_cgs_update_* that returns at once -- w and norm2_out untouched -- when run_flag[0] == 0
 * (run_flag NULL: always runs).  Everything stays on the device and in stream order (CUDA-graph safe). */
int hlv_cgs_needs_pass(const double* c, int rows, const double* norm2, double tol, int* flag_out, hlv_stream_t stream);
int hlv_cgs_update_if_f32 (const float*    V, int64_t ldv, int rows, const double* c, float sign,
                           float* w, int64_t n, double* norm2_out, const int* run_flag,
                           void* ws, size_t ws_bytes, hlv_stream_t stream);
int hlv_cgs_update_if_bf16(const uint16_t* V, int64_t ldv, int rows, const double* c, float sign,
                           float* w, int64_t n, double* norm2_out, const int* run_flag,
                           void* ws, size_t ws_bytes, hlv_stream_t stream);

/* Fused middle pass of two-pass Gram-Schmidt: w -= V^T c_in ; c_out = V w_new ; norm2_out = |w_new|^2,
 * reading V from HBM once (a [rows x tile] slab is staged in shared memory by 2-D tiled TMA boxes of
 * [8 rows x 256 columns] and used for both the update and the projection).  CGS2 = project,
 * update_project, update: 3 passes over V instead of 4.  rows <= hlv_cgs_fused_max_rows(sizeof(elem))
 * (104 fp32 / 208 bf16: two 104 KB slabs per SM); ws must hold rows+1 partial rows:
 * hlv_workspace_bytes(rows + 1).  The launch encodes a CUtensorMap for V on the host
 * (cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: no allocation, no synchronisation), so V must
 * stay 16-byte aligned with ldv*sizeof(elem) a multiple of 16.
 * Replaces: the second/third sweep of the reference's reorthogonalisation loop
 * (Lanczos_Scratch/Discrepancy.ipynb cell 1:44-45 applied twice). */
int hlv_cgs_fused_max_rows(int elem_bytes);
int hlv_cgs_update_project_f32 (const float*    V, int64_t ldv, int rows, const double* c_in,
                                float* w, int64_t n, double* c_out, double* norm2_out,
                                void* ws, size_t ws_bytes, hlv_stream_t stream);
int hlv_cgs_update_project_bf16(const uint16_t* V, int64_t ldv, int rows, const double* c_in,
                                float* w, int64_t n, double* c_out, double* norm2_out,
                                void* ws, size_t ws_bytes, hlv_stream_t stream);

/* ---- (e) multi-GPU: the exchange steps fused into the recurrence kernels over NVLink peer memory -------------
 * The reference has one NCCL site (distributed_scratch.py:8) and otherwise DataParallel; the sharded path here
 *  has three exchange steps per iteration: reduce-scatter of Hv, k-float reductions of alpha /
 * Gram-Schmidt coefficients / |w|^2, all-gather of v_{j+1}.  With a peer context they are not collective launches:
 *   - hlv_x_reduce_scatter_dot_f32 reads its shard of every rank's Hv through peer loads, adds in rank order, writes w
 *     and the alpha partial in the same pass;
 *   - a kernel that finishes a reduction pushes its partial sums into every rank's exchange area from its last CTA,
 *     and the kernel that needs the totals adds the ranks' contributions (rank order, float64) in its prologue;
 *   - hlv_x_normalize_store_f32 writes v_{j+1}'s shard straight into every rank's full-length vector.
 * hlv_peer_ctx: world/rank and, for every rank p, the address of p's exchange area AS MAPPED IN THIS PROCESS
 * (CUDA IPC / symmetric memory; xchg[rank] is the local one).  Each area is hlv_peer_xchg_bytes() bytes, zeroed once
 * by its owner (hlv_peer_xchg_init) before any rank uses it.  One rank per GPU: the waits spin on the device.
 * A NULL context (or world == 1) makes every hlv_x_* call the single-GPU operation.
 * Channels (fixed by convention, one push + one pull per iteration each): */
#define HLV_CH_HV      0   /* flag: this rank's full-length Hv is complete */
#define HLV_CH_ALPHA   1   /* 1 value:  <w, v_j> partial */
#define HLV_CH_C1      2   /* rows values: first projection */
#define HLV_CH_C2      3   /* rows + 1 values: second projection and |w'|^2 */
#define HLV_CH_NORM    4   /* 1 value: |w|^2 partial */
#define HLV_CH_V       5   /* flag: this rank's shard of v_{j+1} has been written into every rank's vector */
typedef struct hlv_peer_ctx {
    int32_t  world, rank;
    uint32_t spin_timeout_ms;            /* 0 = 20 s; a wait that runs out sets the area's error word */
    uint32_t reserved;
    void*    xchg[HLV_MAX_PEERS];
} hlv_peer_ctx;
size_t hlv_peer_xchg_bytes(void);
int    hlv_peer_xchg_init(void* xchg_local, hlv_stream_t stream);
/* error word of the local area (0 = ok, 1 + channel = a wait on that channel timed out); synchronises `stream` */
int    hlv_peer_xchg_error(const void* xchg_local, int* h_error_out, hlv_stream_t stream);
/* flag-only push / wait on a channel (HLV_CH_HV after the HVP, HLV_CH_V before it) */
int    hlv_peer_signal(const hlv_peer_ctx* h_ctx, int channel, hlv_stream_t stream);
int    hlv_peer_wait(const hlv_peer_ctx* h_ctx, int channel, hlv_stream_t stream);

/* w[i] = sum_p hv_p[shard_lo + i] (p = 0..world-1 in order; h_hv[p] = rank p's full-length Hv as mapped here),
 * alpha partial = <w, v> -> alpha_out[0] (local) and pushed on HLV_CH_ALPHA.  Waits for HLV_CH_HV first.
 * hv_multicast (optional): the NVSwitch multicast address of the same buffers -- the sum is then formed IN THE SWITCH
 * (multimem.ld_reduce: one load per element instead of `world`; the reduction order is the switch's, still deterministic).
 * With a NULL context: w = hv_0[shard_lo..], alpha_out = <w, v> (the total).  Replaces reduce_scatter + dot + all_reduce. */
int hlv_x_reduce_scatter_dot_f32(const hlv_peer_ctx* h_ctx, const float* const* h_hv, const float* hv_multicast,
                                 int64_t shard_lo, int64_t n, float* w, const float* v, double* alpha_out,
                                 void* ws, size_t ws_bytes, hlv_stream_t stream);
/* Three-term update folded into the first projection (lanczostrain_hand.py:202, then the first sweep of the
 * reorthogonalisation): alpha = total of HLV_CH_ALPHA (NULL context: alpha[0] as given) and is stored to alpha[0];
 * w -= alpha*vj + beta*vjm1 with torch's rounding sequence (vjm1/beta NULL on the first iteration); c_out = V w_new
 * (local partial; pushed on HLV_CH_C1).  vj / vjm1 are fp32 vectors (rows of an fp32 basis, or the ring). */
int hlv_x_update_project_f32 (const hlv_peer_ctx* h_ctx, const float*    V, int64_t ldv, int rows, float* w, int64_t n,
                              const float* vj, const float* vjm1, double* alpha, const double* beta,
                              double* c_out, void* ws, size_t ws_bytes, hlv_stream_t stream);
int hlv_x_update_project_bf16(const hlv_peer_ctx* h_ctx, const uint16_t* V, int64_t ldv, int rows, float* w, int64_t n,
                              const float* vj, const float* vjm1, double* alpha, const double* beta,
                              double* c_out, void* ws, size_t ws_bytes, hlv_stream_t stream);
/* hlv_lanczos_update_f32 with alpha pulled from HLV_CH_ALPHA and |w|^2 pushed on HLV_CH_NORM (no-reorth runs). */
int hlv_x_lanczos_update_f32(const hlv_peer_ctx* h_ctx, float* w, const float* vj, const float* vjm1,
                             double* alpha, const double* beta, int64_t n, double* norm2_out,
                             void* ws, size_t ws_bytes, hlv_stream_t stream);
/* hlv_cgs_update_project_*: c_in = totals of HLV_CH_C1 (stored back to c_in), c_out / norm2_out local partials pushed
 * together on HLV_CH_C2. */
int hlv_x_cgs_update_project_f32 (const hlv_peer_ctx* h_ctx, const float*    V, int64_t ldv, int rows, double* c_in,
                                  float* w, int64_t n, double* c_out, double* norm2_out,
                                  void* ws, size_t ws_bytes, hlv_stream_t stream);
int hlv_x_cgs_update_project_bf16(const hlv_peer_ctx* h_ctx, const uint16_t* V, int64_t ldv, int rows, double* c_in,
                                  float* w, int64_t n, double* c_out, double* norm2_out,
                                  void* ws, size_t ws_bytes, hlv_stream_t stream);
/* hlv_cgs_update_*: c = the first `rows` totals of HLV_CH_C2 (stored back to c), |w|^2 pushed on HLV_CH_NORM. */
int hlv_x_cgs_update_f32 (const hlv_peer_ctx* h_ctx, const float*    V, int64_t ldv, int rows, double* c,
                          float* w, int64_t n, double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream);
int hlv_x_cgs_update_bf16(const hlv_peer_ctx* h_ctx, const uint16_t* V, int64_t ldv, int rows, double* c,
                          float* w, int64_t n, double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream);
/* hlv_normalize_store_f32 with norm2 = total of HLV_CH_NORM (stored back to norm2[0]); additionally the normalised
 * shard is written to h_v_full[p] + shard_lo for every rank p (h_v_full[p] = rank p's full-length vector as mapped
 * here; NULL list = no peer writes; v_multicast, optional: ONE multimem.st per element that the switch replicates to
 * every rank instead of `world` peer stores) and HLV_CH_V is pushed when all of it is on its way.  v_out / row_bf16 as in
 * hlv_normalize_store_f32; with everything NULL only beta is produced (and nothing is pushed). */
int hlv_x_normalize_store_f32(const hlv_peer_ctx* h_ctx, const float* w, double* norm2, int64_t n,
                              double* beta_out, float* v_out, uint16_t* row_bf16,
                              float* const* h_v_full, float* v_multicast, int64_t shard_lo,
                              double breakdown_tol, int* breakdown_iter, int iter,
                              void* ws, size_t ws_bytes, hlv_stream_t stream);

/* ---- low-rank gradient adjustment: drop-in for vector_adjust.cu:2-15 -------- */
/* adjusted[x] += sum_i (1/eig[i] - 1/(eig[i]+delta)) * (grad . V_i) * V[i*ldv + x].
 * Same argument meaning and order as the reference kernel (+ ldv, scratch, stream);
 * two streaming passes over V instead of the reference's O(k*n^2) loads.
 * coef_scratch: device doubles, >= num_eigenvalues. */
int hlv_vector_adjust_f32(const float* grad_vector, const float* V, const float* eigvals,
                          float* adjusted_grad_vector, int num_eigenvalues, int64_t vec_len,
                          float delta, int64_t ldv, double* coef_scratch,
                          void* ws, size_t ws_bytes, hlv_stream_t stream);

/* ---- Ritz vectors: out[r*ldo + x] = sum_i Y[i*ldy + r] * Q[i*ldq + x] --------
 * (V = eigvects.t() @ Q, gpt2_hessian_cpu.py:217 / lanczostrain_hand.py:210)
 * Y is a DEVICE fp32 m x ldy matrix whose columns are eigenvectors of T; the
 * first `nvec` columns are materialised. */
int hlv_ritz_vectors_f32 (const float*    Q, int64_t ldq, int m, const float* Y, int ldy, int nvec,
                          float* out, int64_t ldo, int64_t n, hlv_stream_t stream);
int hlv_ritz_vectors_bf16(const uint16_t* Q, int64_t ldq, int m, const float* Y, int ldy, int nvec,
                          float* out, int64_t ldo, int64_t n, hlv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HLV_H_ */
