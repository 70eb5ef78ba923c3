// libhlv.so -- kernel (c'), the middle pass of CGS2 fused:   w' = w - V^T c1 ;  c2 = V w' ;  |w'|^2
// reading V from HBM ONCE instead of twice.
//
// Two-pass classical Gram-Schmidt is project, update, project, update = 4 streaming passes over the
// basis.  The 2nd and 3rd touch the same data in the same order, and the 3rd needs nothing global from
// the 2nd (w' is complete per column as soon as all rows of that column have been applied).  On B200 a
// CTA can hold a whole [rows x TW] column slab of the basis in shared memory (104 fp32 rows x 256 columns
// = 104 KB; two such CTAs per SM), so: apply the update from the slab as its rows land (pass A), then
// project the finished w' tile against the SAME slab (pass B).  CGS2 becomes 3 passes:
// (3*j*s + 20)*n bytes instead of (4*j*s + 24)*n.
//
// Data movement is 2-D tiled TMA (cp.async.bulk.tensor.2d, SASS UTMALDG): the basis is described to the
// TMA unit as a [rows x n] tensor with row pitch ldv, and the producer warp asks for one [8 rows x 256
// columns] box per instruction -- 8 KB per request.  (The first version issued one 1-D bulk copy per row;
// at 256-column tiles that is a 1 KB request every ~50 cycles per SM, and the request rate, not HBM,
// bounded the kernel at 4.9 TB/s.)  Rows past `rows` and columns past n are out of bounds for the tensor
// map and arrive as zeros, so neither the last row group nor the ragged last tile needs special code.
// Completion is tracked by one mbarrier per 8-row group; a group's slots are handed back to the producer
// the moment pass B has consumed them, so the next tile's boxes are in flight while this tile is still
// being projected.  8 consumer warps: pass A "thread owns a 16-byte chunk of a row slice", pass B "warp
// owns row groups" (fixed-order reductions).
#include <cuda.h>        // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved through the runtime
#include <limits.h>
#include <stdlib.h>

#include <mutex>

#include "hlv_peer.cuh"

namespace hlv {

constexpr int kFusedConsumers = 256;
constexpr int kFusedThreads = kFusedConsumers + 32;      // + one producer warp
constexpr int kGroup = 8;                                // rows per mbarrier = rows per TMA box
constexpr int kBoxCols = 256;                            // columns per TMA box (the hardware maximum per dimension)
// Two CTAs per SM, each with a <= 104 KB slab: while one CTA waits for its tile to land, the other is in its
// compute passes, so HBM stays busy (a single 208 KB slab cannot overlap a tile's load with its own compute).
constexpr int kFusedCtasPerSm = 2;
constexpr int kFusedSlabBytes = 104 * 1024;

// ---- mbarrier / TMA PTX ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one [kGroup x kBoxCols] box of the tensor at column x, row y -> shared memory, completing on `bar`
__device__ __forceinline__ void tma_box_g2s(void* smem_dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kFusedConsumers) : "memory"); }

// 16 bytes of a slab row -> floats (4 fp32 or 8 bf16)
template <typename BT> struct Chunk16;
template <> struct Chunk16<float> {
    static constexpr int kElems = 4;
    static __device__ __forceinline__ void load(const float* p, float (&x)[8]) {
        float4 v = *reinterpret_cast<const float4*>(p);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
};
template <> struct Chunk16<uint16_t> {
    static constexpr int kElems = 8;
    static __device__ __forceinline__ void load(const uint16_t* p, float (&x)[8]) {
        uint4 v = *reinterpret_cast<const uint4*>(p);
        x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x); x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
        x[4] = bf16_lo(v.z); x[5] = bf16_hi(v.z); x[6] = bf16_lo(v.w); x[7] = bf16_hi(v.w);
    }
};

// CPT consecutive floats of w / s_w as one access when the tile is full
template <int CPT> __device__ __forceinline__ void ld_f32(const float* p, float (&v)[CPT]) {
    if (CPT == 2) { float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1 % CPT] = t.y; }
    else if (CPT == 4) { float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1 % CPT] = t.y; v[2 % CPT] = t.z; v[3 % CPT] = t.w; }
    else if (CPT == 8) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1 % CPT] = a.y; v[2 % CPT] = a.z; v[3 % CPT] = a.w; v[4 % CPT] = b.x; v[5 % CPT] = b.y; v[6 % CPT] = b.z; v[7 % CPT] = b.w;
    } else { v[0] = p[0]; }
}
template <int CPT> __device__ __forceinline__ void st_f32(float* p, const float (&v)[CPT]) {
    if (CPT == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1 % CPT]);
    else if (CPT == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1 % CPT], v[2 % CPT], v[3 % CPT]);
    else if (CPT == 8) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1 % CPT], v[2 % CPT], v[3 % CPT]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4 % CPT], v[5 % CPT], v[6 % CPT], v[7 % CPT]);
    } else p[0] = v[0];
}

struct FusedSmem {
    size_t slab, s_w, s_part, s_c, s_acc, bars, total;
};
// Pass A works on 16-byte chunks: a row of the tile is tw*elem_bytes/16 chunks, shared by 256 threads as
// `slices` = 256 / chunks row slices (1 for the widest tile, 4 for the 256-column fp32 tile).
__host__ __device__ inline int fused_slices(int tw, int elem_bytes) { return kFusedConsumers / (tw * elem_bytes / 16); }
// Slab = [group][box][8 rows][256 columns]: box (g, b) holds rows 8g..8g+7, columns 256b..256b+255 of the tile.
__host__ __device__ inline FusedSmem fused_layout(int rows, int tw, int elem_bytes) {
    FusedSmem L;
    const int ngroups = (rows + kGroup - 1) / kGroup;
    const int slices = fused_slices(tw, elem_bytes);
    L.slab = 0;
    L.s_w = (size_t)ngroups * kGroup * tw * elem_bytes;
    L.s_part = L.s_w + (size_t)tw * 4;                       // [slices][tw] partial updates (only when slices > 1)
    L.s_c = L.s_part + (slices > 1 ? (size_t)slices * tw * 4 : 0);
    L.s_acc = L.s_c + (size_t)ngroups * kGroup * 4;        // c and the accumulators are padded to whole groups
    L.bars = L.s_acc + (size_t)ngroups * kGroup * 8;
    L.total = L.bars + (size_t)ngroups * 16;
    return L;
}

template <typename BT, int CPT>
__global__ void __launch_bounds__(kFusedThreads, kFusedCtasPerSm)
cgs_update_project_kernel(const __grid_constant__ CUtensorMap tmap, int rows, double* c_in,
                          float* __restrict__ w, int64_t n, double* partials, unsigned* counter,
                          double* c_out, double* norm2_out, const __grid_constant__ PeerView pv) {
    constexpr int TW = kFusedConsumers * CPT;               // tile width = CPT boxes
    constexpr int kBoxElems = kGroup * kBoxCols;
    constexpr uint32_t kBoxBytes = kBoxElems * sizeof(BT);
    extern __shared__ __align__(128) unsigned char smem[];
    const FusedSmem L = fused_layout(rows, TW, (int)sizeof(BT));
    BT* slab = reinterpret_cast<BT*>(smem + L.slab);
    float* s_w = reinterpret_cast<float*>(smem + L.s_w);
    float* s_part = reinterpret_cast<float*>(smem + L.s_part);
    float* s_c = reinterpret_cast<float*>(smem + L.s_c);
    double* s_acc = reinterpret_cast<double*>(smem + L.s_acc);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
    const int ngroups = (rows + kGroup - 1) / kGroup;
    uint64_t* empty = full + ngroups;
    __shared__ double s_warp[kWarps];
    __shared__ unsigned s_ticket;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int g = 0; g < ngroups; ++g) {
            mbar_init(&full[g], CPT);                // one arrive.expect_tx per box of the group (+ the boxes' bytes)
            mbar_init(&empty[g], 1);                 // the owning consumer warp's release after pass B
        }
        mbar_fence_init();
    }
    if (pv.world > 1) {                                 // c_in = rank-ordered totals of the first projection
        const bool keep = blockIdx.x == 0;
        peer_pull(pv, HLV_CH_C1, rows, [&](int i, double t) { s_c[i] = -(float)t; if (keep) c_in[i] = t; });
        for (int i = tid; i < ngroups * kGroup; i += kFusedThreads) {
            if (i >= rows) s_c[i] = 0.0f;
            s_acc[i] = 0.0;
        }
    } else {
        for (int i = tid; i < ngroups * kGroup; i += kFusedThreads) {
            s_c[i] = i < rows ? -(float)c_in[i] : 0.0f;
            s_acc[i] = 0.0;
        }
    }
    __syncthreads();

    const int64_t ntiles = (n + TW - 1) / TW;
    float nrm = 0.0f;

    if (warp == kWarps) {
        // ===== producer warp: one lane per (group, box); a lane waits for its group's slots, announces its
        //       box's bytes on the group's barrier and issues the box =====
        const int nitems = ngroups * CPT;
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            for (int item = lane; item < nitems; item += 32) {
                const int g = item / CPT, b = item - g * CPT;
                mbar_wait(&empty[g], (it & 1u) ^ 1u);              // slot group free (passes at once on the first tile)
                mbar_arrive_expect_tx(&full[g], kBoxBytes);
                tma_box_g2s(slab + (size_t)item * kBoxElems, &tmap, (int)(tile * TW) + b * kBoxCols, g * kGroup, &full[g]);
            }
        }
    } else {
        // ===== consumers =====
        // Pass A mapping: a thread always works on one 16-byte chunk of a row (E elements, one LDS.128 per E
        // FMAs).  A tile row has NQ chunks; when the tile is narrow (NQ < 256) the 256 threads split the row
        // groups into S = 256/NQ slices (slice s takes groups s, s+S, ...), and the S partial updates are
        // summed in slice order through shared memory -- same instruction count per byte for every width.
        constexpr int E = Chunk16<BT>::kElems;                  // elements per 16-byte chunk
        constexpr int NQ = TW / E;                              // chunks per tile row
        constexpr int S = kFusedConsumers / NQ;                 // row slices
        static_assert(NQ * S == kFusedConsumers && (S > 1 || E == CPT), "tile width / slice mapping");
        const int q = tid % NQ, slice = tid / NQ;
        // element offset of this thread's chunk inside a group: box (q*E / 256), column (q*E % 256)
        const int a_off = ((q * E) / kBoxCols) * kBoxElems + ((q * E) % kBoxCols);
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int64_t x0 = tile * TW;
            const int valid = (int)((n - x0 < TW) ? (n - x0) : TW);
            const bool whole = valid == TW;
            float acc[E];
            if (slice == 0 && whole) {
                ld_f32<E>(w + x0 + q * E, acc);
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int x = q * E + e;
                    acc[e] = (slice == 0 && x < valid) ? w[x0 + x] : 0.0f;
                }
            }
            // ---- pass A: w' = w - sum_i c1[i] V[i, tile]  (rows >= `rows` are zeros with zero coefficients) ----
            for (int g = slice; g < ngroups; g += S) {
                mbar_wait(&full[g], it & 1u);
                const int i0 = g * kGroup;
                const float4 ca = *reinterpret_cast<const float4*>(s_c + i0);       // 8 coefficients, 2 broadcasts
                const float4 cb = *reinterpret_cast<const float4*>(s_c + i0 + 4);
                const float cg[kGroup] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
                const BT* col = slab + (size_t)g * (CPT * kBoxElems) + a_off;
                float v[kGroup][8];
#pragma unroll
                for (int r = 0; r < kGroup; ++r) Chunk16<BT>::load(col + r * kBoxCols, v[r]);   // all loads first ...
#pragma unroll
                for (int r = 0; r < kGroup; ++r) {                                              // ... then the FMAs
#pragma unroll
                    for (int e = 0; e < E; ++e) acc[e] = fmaf(cg[r], v[r][e], acc[e]);
                }
            }
            if constexpr (S == 1) {
                consumer_bar();                                    // nobody still reads s_w from the previous pass B
                if (whole) {
                    st_f32<E>(w + x0 + q * E, acc);
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        if (q * E + e < valid) w[x0 + q * E + e] = acc[e];
                }
                st_f32<E>(s_w + q * E, acc);                       // columns >= valid hold 0
#pragma unroll
                for (int e = 0; e < E; ++e) nrm = fmaf(acc[e], acc[e], nrm);
            } else {
                st_f32<E>(s_part + slice * TW + q * E, acc);
                consumer_bar();                                    // partials visible; previous pass B is over too
                float fin[CPT];
                ld_f32<CPT>(s_part + tid * CPT, fin);
#pragma unroll
                for (int sl = 1; sl < S; ++sl) {                   // fixed slice order
                    float t[CPT];
                    ld_f32<CPT>(s_part + sl * TW + tid * CPT, t);
#pragma unroll
                    for (int c = 0; c < CPT; ++c) fin[c] += t[c];
                }
                if (whole) {
                    st_f32<CPT>(w + x0 + tid * CPT, fin);
                } else {
#pragma unroll
                    for (int c = 0; c < CPT; ++c)
                        if (tid * CPT + c < valid) w[x0 + tid * CPT + c] = fin[c];
                }
                st_f32<CPT>(s_w + tid * CPT, fin);                 // columns >= valid hold 0
#pragma unroll
                for (int c = 0; c < CPT; ++c) nrm = fmaf(fin[c], fin[c], nrm);
            }
            consumer_bar();                                        // w' tile complete in shared memory
            // ---- pass B: c2[i] += <V[i, tile], w'>.  Warp (g mod 8) owns row group g: its lanes keep their
            //      columns of w' in registers, run 8 independent row accumulators, and reduce all 8 rows with
            //      ONE 9-shuffle butterfly; then the group's slots go back to the producer. ----
            constexpr int NCH = TW / (32 * E);                      // chunks per lane per row
            float wr[NCH][E];
            int b_off[NCH];
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                const int x = (k * 32 + lane) * E;
                ld_f32<E>(s_w + x, wr[k]);
                b_off[k] = (x / kBoxCols) * kBoxElems + (x % kBoxCols);
            }
            const int my_row = warp_sum8_row(lane);
            for (int g = warp; g < ngroups; g += kWarps) {
                const BT* grp = slab + (size_t)g * (CPT * kBoxElems);
                float p[kGroup];
#pragma unroll
                for (int r = 0; r < kGroup; ++r) {
                    p[r] = 0.0f;
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        float v[8];
                        Chunk16<BT>::load(grp + b_off[k] + r * kBoxCols, v);
#pragma unroll
                        for (int e = 0; e < E; ++e) p[r] = fmaf(v[e], wr[k][e], p[r]);
                    }
                }
                const float tot = warp_sum8(p, lane);
                if ((lane & 3) == 0) s_acc[g * kGroup + my_row] += (double)tot;
                __syncwarp();                                       // every lane is done reading the group's rows
                if (lane == 0) mbar_arrive(&empty[g]);              // hand the group's slots back to the producer
            }
        }
    }
    // ===== epilogue: per-CTA partials, then the last CTA reduces across CTAs in index order =====
    double t = warp_sum((double)nrm);
    if (warp < kWarps && lane == 0) s_warp[warp] = t;
    __syncthreads();
    for (int r = tid; r < rows; r += kFusedThreads) partials[(size_t)r * kMaxCtas + blockIdx.x] = s_acc[r];
    if (tid == 0) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < kWarps; ++i) s += s_warp[i];
        partials[(size_t)rows * kMaxCtas + blockIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicInc(counter, gridDim.x - 1);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    if (warp < kWarps) {
        const int nblk = gridDim.x;
        for (int r = warp; r <= rows; r += kWarps) {
            const double* p = partials + (size_t)r * kMaxCtas;
            double s = 0.0;
            for (int b = lane; b < nblk; b += 32) s += __ldcg(p + b);
            s = warp_sum(s);
            if (lane == 0) {
                if (r < rows) c_out[r] = s; else norm2_out[0] = s;
            }
        }
    }
    if (pv.world > 1) {                                 // push (c2[0..rows), |w'|^2) as ONE message on HLV_CH_C2
        __syncthreads();                                // every warp is done with the partials: row 0's region is free
        double* stage = partials;
        for (int i = tid; i <= rows; i += kFusedThreads) stage[i] = i < rows ? c_out[i] : norm2_out[0];
        __syncthreads();
        peer_push(pv, HLV_CH_C2, stage, rows + 1);
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn lookup_encode_tiled() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
}
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = lookup_encode_tiled();       // resolved once (thread-safe static)
    return fn;
}

// The basis as the TMA unit sees it: [rows x n] elements, row pitch ldv, fetched in [8 x 256] boxes, zeros outside.
template <typename BT>
static int make_basis_map(CUtensorMap* map, const BT* V, int64_t ldv, int rows, int64_t n, const char* name) {
    EncodeTiledFn enc = encode_tiled();
    HLV_REQUIRE(enc != nullptr, HLV_ERR_NO_DEVICE, "%s: cuTensorMapEncodeTiled is not available from this driver", name);
    const cuuint64_t gdim[2] = {(cuuint64_t)(n > 0 ? n : 1), (cuuint64_t)rows};   // n == 0: no tile is ever requested
    const cuuint64_t gstride[1] = {(cuuint64_t)ldv * sizeof(BT)};
    const cuuint32_t box[2] = {(cuuint32_t)kBoxCols, (cuuint32_t)kGroup};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = enc(map, sizeof(BT) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                           const_cast<BT*>(V), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HLV_REQUIRE(r == CUDA_SUCCESS, HLV_ERR_ARG, "%s: cuTensorMapEncodeTiled failed (CUresult %d) for rows=%d n=%lld ldv=%lld",
                name, (int)r, rows, (long long)n, (long long)ldv);
    return HLV_OK;
}

// Widest tile (columns) whose slab fits; 0 if even the narrowest does not.
// Slab budget per CTA.  Default 104 KB = two CTAs per SM; HLV_FUSED_SLAB_KB (tuning knob, read once) can raise
// it to ~208 KB = one CTA per SM with a tile twice as wide.
static size_t slab_budget() {
    static size_t budget = 0;
    if (budget == 0) {
        const char* e = getenv("HLV_FUSED_SLAB_KB");
        long kb = e ? atol(e) : 0;
        budget = (kb >= 16 && kb <= 208) ? (size_t)kb * 1024 : (size_t)kFusedSlabBytes;
    }
    return budget;
}

// Two CTAs per SM is what hides a tile's load behind the other CTA's compute passes, so the tile is the widest one whose
// WHOLE shared-memory layout (slab + w tile + slice partials + coefficients + barriers) fits twice into the SM's 227 KB
// (1 KB per CTA is reserved by the system).  Round 1 budgeted the slab alone: bf16 rows at 512-column tiles came to 118 KB,
// one CTA per SM, and ncu showed the refill bubble (57% of DRAM peak).
constexpr size_t kFusedTotalPerCta = 112 * 1024;
template <typename BT>
static int pick_cpt(int rows) {
    const int max_cpt = sizeof(BT) == 4 ? 4 : 8;          // one 16-byte chunk per thread per row at the widest
    const int rows_pad = (rows + kGroup - 1) / kGroup * kGroup;          // the slab holds whole 8-row boxes
    const bool two_ctas = slab_budget() == (size_t)kFusedSlabBytes;      // default policy (no HLV_FUSED_SLAB_KB override)
    for (int cpt = max_cpt; cpt >= 1; cpt >>= 1) {
        if ((size_t)rows_pad * kFusedConsumers * cpt * sizeof(BT) > slab_budget()) continue;
        if (two_ctas && cpt > 1 && fused_layout(rows, kFusedConsumers * cpt, (int)sizeof(BT)).total > kFusedTotalPerCta) continue;
        return cpt;
    }
    return 0;
}

// The host encodes one CUtensorMap per (basis pointer, pitch, rows, n, element type); a Lanczos run asks for the same
// ~100 maps on every restart, so the last few hundred are kept (direct-mapped, keyed by the arguments).
struct MapKey {
    const void* V; int64_t ldv, n; int rows, elem;
    bool operator==(const MapKey& o) const { return V == o.V && ldv == o.ldv && n == o.n && rows == o.rows && elem == o.elem; }
};
struct MapEntry { MapKey key; CUtensorMap map; bool used; };
static MapEntry g_maps[512];
static std::mutex g_maps_mu;

template <typename BT>
static int basis_map(CUtensorMap* out, const BT* V, int64_t ldv, int rows, int64_t n, const char* name) {
    const MapKey key{V, ldv, n, rows, (int)sizeof(BT)};
    const size_t h = (reinterpret_cast<uintptr_t>(V) >> 8) * 1315423911u + (size_t)rows * 2654435761u + (size_t)n * 97 + (size_t)ldv * 31 + sizeof(BT);
    MapEntry& e = g_maps[h % 512];
    {
        std::lock_guard<std::mutex> lock(g_maps_mu);
        if (e.used && e.key == key) { *out = e.map; return HLV_OK; }
    }
    const int rc = make_basis_map<BT>(out, V, ldv, rows, n, name);
    if (rc != HLV_OK) return rc;
    std::lock_guard<std::mutex> lock(g_maps_mu);
    e.key = key; e.map = *out; e.used = true;
    return HLV_OK;
}

template <typename BT, int CPT>
static int launch_fused(const PeerView& pv, const BT* V, int64_t ldv, int rows, double* c_in, float* w, int64_t n, const Workspace& ws,
                        double* c_out, double* norm2_out, cudaStream_t stream, const char* name) {
    const size_t smem = fused_layout(rows, kFusedConsumers * CPT, (int)sizeof(BT)).total;
    const void* fn = reinterpret_cast<const void*>(cgs_update_project_kernel<BT, CPT>);
    cudaError_t e = ensure_dynamic_smem(fn, smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(update_project)");
    const int64_t tw = kFusedConsumers * CPT;
    const int grid = persistent_grid((n + tw - 1) / tw, cached_occupancy(fn, kFusedThreads, smem));
    CUtensorMap map;
    const int rc = basis_map<BT>(&map, V, ldv, rows, n, name);
    if (rc != HLV_OK) return rc;
    cgs_update_project_kernel<BT, CPT><<<grid, kFusedThreads, smem, stream>>>(map, rows, c_in, w, n, ws.partials, ws.counters,
                                                                              c_out, norm2_out, pv);
    HLV_LAUNCH_CHECK(name);
    return HLV_OK;
}

template <typename BT>
static int update_project(const char* name, const hlv_peer_ctx* h_ctx, const BT* V, int64_t ldv, int rows, double* c_in, float* w, int64_t n,
                          double* c_out, double* norm2_out, void* ws_raw, size_t ws_bytes, cudaStream_t stream) {
    HLV_REQUIRE(V && w && c_in && c_out && norm2_out && n >= 0 && rows >= 1, HLV_ERR_ARG, "%s: bad argument", name);
    const int prc = check_peer_ctx(h_ctx, name);
    if (prc != HLV_OK) return prc;
    const PeerView pv = make_peer_view(h_ctx);
    HLV_REQUIRE(ldv >= n, HLV_ERR_ARG, "%s: ldv=%lld < n=%lld", name, (long long)ldv, (long long)n);
    HLV_REQUIRE(aligned16(V) && aligned16(w) && ((ldv * (int64_t)sizeof(BT)) & 15) == 0, HLV_ERR_ALIGN,
                "%s: V, w must be 16-byte aligned and ldv*sizeof(elem) a multiple of 16", name);
    HLV_REQUIRE(n <= (int64_t)INT32_MAX - 4096, HLV_ERR_ARG,
                "%s: n=%lld exceeds the 32-bit TMA tile coordinates; shard the vector or use project+update", name, (long long)n);
    const int cpt = pick_cpt<BT>(rows);
    HLV_REQUIRE(cpt > 0, HLV_ERR_ARG, "%s: rows=%d exceeds the fused kernel's shared-memory slab (max %d); use project+update",
                name, rows, hlv_cgs_fused_max_rows((int)sizeof(BT)));
    Workspace ws;
    HLV_REQUIRE(carve_workspace(ws_raw, ws_bytes, rows + 1, &ws), HLV_ERR_WORKSPACE,
                "%s: workspace too small for %d+1 rows (need %zu bytes)", name, rows, workspace_bytes(rows + 1));
    HLV_REQUIRE(sm_count() > 0, HLV_ERR_NO_DEVICE, "%s: no CUDA device", name);
    switch (cpt) {
        case 8:
            if constexpr (sizeof(BT) == 2) return launch_fused<BT, 8>(pv, V, ldv, rows, c_in, w, n, ws, c_out, norm2_out, stream, name);
        case 4: return launch_fused<BT, 4>(pv, V, ldv, rows, c_in, w, n, ws, c_out, norm2_out, stream, name);
        case 2: return launch_fused<BT, 2>(pv, V, ldv, rows, c_in, w, n, ws, c_out, norm2_out, stream, name);
        default: return launch_fused<BT, 1>(pv, V, ldv, rows, c_in, w, n, ws, c_out, norm2_out, stream, name);
    }
}

}  // namespace hlv

using namespace hlv;

extern "C" {

int hlv_cgs_fused_max_rows(int elem_bytes) {
    if (elem_bytes != 2 && elem_bytes != 4) return 0;
    return (int)(slab_budget() / (size_t)(kFusedConsumers * elem_bytes)) / kGroup * kGroup;
}

int hlv_cgs_update_project_f32(const float* V, int64_t ldv, int rows, const double* c_in, float* w, int64_t n,
                               double* c_out, double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update_project<float>("hlv_cgs_update_project_f32", nullptr, V, ldv, rows, const_cast<double*>(c_in), w, n, c_out, norm2_out, ws, ws_bytes,
                                 static_cast<cudaStream_t>(stream));
}
int hlv_cgs_update_project_bf16(const uint16_t* V, int64_t ldv, int rows, const double* c_in, float* w, int64_t n,
                                double* c_out, double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update_project<uint16_t>("hlv_cgs_update_project_bf16", nullptr, V, ldv, rows, const_cast<double*>(c_in), w, n, c_out, norm2_out, ws, ws_bytes,
                                    static_cast<cudaStream_t>(stream));
}
int hlv_x_cgs_update_project_f32(const hlv_peer_ctx* h_ctx, const float* V, int64_t ldv, int rows, double* c_in, float* w, int64_t n,
                                 double* c_out, double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update_project<float>("hlv_x_cgs_update_project_f32", h_ctx, V, ldv, rows, c_in, w, n, c_out, norm2_out, ws, ws_bytes,
                                 static_cast<cudaStream_t>(stream));
}
int hlv_x_cgs_update_project_bf16(const hlv_peer_ctx* h_ctx, const uint16_t* V, int64_t ldv, int rows, double* c_in, float* w, int64_t n,
                                  double* c_out, double* norm2_out, void* ws, size_t ws_bytes, hlv_stream_t stream) {
    return update_project<uint16_t>("hlv_x_cgs_update_project_bf16", h_ctx, V, ldv, rows, c_in, w, n, c_out, norm2_out, ws, ws_bytes,
                                    static_cast<cudaStream_t>(stream));
}

}  // extern "C"
