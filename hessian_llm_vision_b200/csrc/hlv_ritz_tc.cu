// libhlv.so -- Ritz-vector materialisation  V = Y^T Q  (gpt2_hessian_cpu.py:217, lanczostrain_hand.py:210) as ONE pass
// over the basis on the 5th-generation tensor cores.
//
// This is the one real GEMM of the path: out[r, x] = sum_i Y[i, r] * Q[i, x], i < m <= 128,
// r < nvec <= 112, x < n ~ 1.2e8: 2*m*nvec*n = 2.5 TFLOP at m = nvec = 100 against 99 GB of HBM traffic (read Q once,
// write V once) -- 25 flop/byte.  The CUDA-core kernel (hlv_cgs.cu) produces 8 output vectors per pass and therefore
// re-reads Q ceil(nvec/8) times (13x at nvec = 100: 695 GB).  Here a CTA walks over [m x 128]-column tiles of Q:
//
//   D[x, r] (128 TMEM lanes x N columns, fp32)  +=  A[x, i] (= Q tile, MN-major in shared memory)  *  B[i, r] (= Y, K-major)
//
// issued as tcgen05.mma kind::tf32 with M = 128, N = roundup(nvec, 16), K = 8 per instruction.  fp32 accuracy is kept
// with the 3xTF32 split: q = q_hi + q_lo, y = y_hi + y_lo (hi = the top 19 bits, exactly representable in tf32; lo = the
// remainder, itself truncated to tf32), D += q_hi*y_hi + q_lo*y_hi + q_hi*y_lo -- error ~2^-21 per product, fp32
// accumulation in TMEM.
//
// Warp roles (320 threads, one CTA per SM, persistent over tiles):
//   warp 0      TMA producer: one [16 rows x 128 columns] box (8 KB, row-major as in HBM) per stage into a 3-deep ring;
//               rows >= m arrive as zeros
//   warps 2-5   transpose + split: the contraction runs over the ROWS of the tile, so the A operand must have the basis-row
//               index as its K dimension.  Thread t owns column x of the tile = TMEM lane x: it reads its 16 values of the
//               stage from shared memory (conflict-free scalar loads), splits them into q_hi / q_lo and writes both with
//               one tcgen05.st each into the A-operand ring IN TENSOR MEMORY (lane = x, 16 columns = the stage's rows) --
//               the transpose is free, and the MMA takes A from TMEM, which halves the shared-memory traffic per tile.
//               (The hardware's own transposing path -- an MN-major tf32 shared-memory descriptor over the TMA-swizzled
//               tile -- returned zeros on this part in every descriptor variant tried.)
//   warp 1      MMA issuer (one elected lane): 3 tcgen05.mma (A from TMEM, B = Y from shared memory) per 8-row K atom,
//               tcgen05.commit frees the operand stage for the split warps; after the last atom of a tile a commit hands the
//               accumulator to the epilogue.  Also owns the TMEM allocation (512 columns: two 128-column accumulators, so
//               tile t+1 is multiplied while tile t drains, + 8 operand stages of 32 columns)
//   warps 6-9   epilogue: tcgen05.ld 32x32b (lane = column x of the tile), one coalesced 128-byte store per output row
// Y (hi and lo, zero padded) is staged once per CTA in the same K-major no-swizzle core-matrix layout.
// Tried and dropped (no gain, profiles/README.md): a second group of split warps and of epilogue warps (18 warps: 21.2 ms per
// launch under ncu against 19.9 ms for this 10-warp version), the operand in shared memory (SS form: within 3%).
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>

#include "hlv_common.cuh"

namespace hlv {

constexpr int kTcTileCols = 128;                 // UMMA M
constexpr int kTcStageRows = 16;                 // two K atoms of 8 rows per pipeline stage
constexpr int kTcStages = 8;                     // (hi, lo) operand stages IN TENSOR MEMORY between the split warps and the MMA issuer
constexpr int kTcMaxRawStages = 12;              // raw tiles between the TMA producer and the split warps: as many as fit
                                                 // (HBM latency x 44 GB/s per SM wants >= 48 KB of loads in flight per SM)
constexpr int kTcStageBytes = kTcStageRows * kTcTileCols * 4;        // 8 KB (raw) / 8 KB (hi) + 8 KB (lo)
constexpr int kTcThreads = 320;
constexpr int kTcMaxN = 112;
constexpr int kTcMaxM = 128;
constexpr int kTcTmemCols = 512;                 // two accumulators of up to 128 columns + the A-operand ring
constexpr int kTcTmemA = 256;                    // first column of the A ring: stage s = [hi: 16 columns][lo: 16 columns]

// ---- PTX helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "TC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra TC_DONE;\n"
        "bra TC_WAIT;\n"
        "TC_DONE:\n"
        "}" ::"r"(tc_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_tma_box(void* dst, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(tc_smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {          // arrives on `bar` when every MMA issued so far is done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (lane = row of A, consecutive columns = K), B from shared memory
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 consecutive columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
                   "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane (the caller waits: tcgen05.wait::ld)
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets (16-byte units),
// version 1 (Blackwell), layout type in bits [61,64): 0 = no swizzle, 2 = 128-byte swizzle.
__device__ __forceinline__ uint64_t tc_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32: D = F32 (bits 4-5), A = B = TF32 (bits 7-9, 10-12),
// both operands K-major (bits 15, 16 clear), N >> 3 in bits 17-22, M >> 4 in bits 24-28 (M = 128).
__host__ __device__ inline uint32_t tc_idesc(int n_cols) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

struct RitzTcSmem {
    size_t raw, a_hi, a_lo, b_hi, b_lo, bars, tmem_slot, total;
};
__host__ __device__ inline RitzTcSmem ritz_tc_layout(int kchunks_pad, int n_cols, int raw_stages, bool a_tmem) {
    RitzTcSmem L;
    L.raw = 0;
    L.a_hi = (size_t)raw_stages * kTcStageBytes;                         // A operand stages in shared memory (SS form) ...
    L.a_lo = L.a_hi + (a_tmem ? 0 : (size_t)kTcStages * kTcStageBytes);  // ... or none: the operand lives in tensor memory
    L.b_hi = L.a_lo + (a_tmem ? 0 : (size_t)kTcStages * kTcStageBytes);
    const size_t b_bytes = (size_t)kchunks_pad * n_cols * 32;            // per K chunk: [2 halves][n_cols/8][8][4 floats]
    L.b_lo = L.b_hi + b_bytes;
    L.bars = L.b_lo + b_bytes;
    L.tmem_slot = L.bars + (2 * kTcMaxRawStages + 2 * kTcStages + 4) * sizeof(uint64_t);
    L.total = L.tmem_slot + 16;
    return L;
}

template <bool A_TMEM>
__global__ void __launch_bounds__(kTcThreads, 1)
ritz_vectors_tc_kernel(const __grid_constant__ CUtensorMap tmap, int m, const float* __restrict__ Y, int ldy, int v0, int nvec,
                       int n_cols, float* __restrict__ out, int64_t ldo, int64_t ntiles, int raw_stages) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // generous alignment of the carve-up (TMA destinations need 128 bytes, UMMA core matrices 16)
    unsigned char* smem = smem_raw + ((1024u - (tc_smem_u32(smem_raw) & 1023u)) & 1023u);
    const int kchunks = (m + 7) / 8;
    const int nstages_per_tile = (kchunks + 1) / 2;
    const int kchunks_pad = 2 * nstages_per_tile;
    const RitzTcSmem L = ritz_tc_layout(kchunks_pad, n_cols, raw_stages, A_TMEM);
    unsigned char* raw = smem + L.raw;
    unsigned char* a_hi = smem + L.a_hi;
    unsigned char* a_lo = smem + L.a_lo;
    float* b_hi = reinterpret_cast<float*>(smem + L.b_hi);
    float* b_lo = reinterpret_cast<float*>(smem + L.b_lo);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);       // TMA -> split          [raw_stages]
    uint64_t* raw_free = full + kTcMaxRawStages;                       // split -> TMA          [raw_stages]
    uint64_t* ready = raw_free + kTcMaxRawStages;                      // split -> MMA          [kTcStages]
    uint64_t* empty = ready + kTcStages;                               // MMA -> split          [kTcStages]
    uint64_t* tmem_full = empty + kTcStages;                           // MMA -> epilogue   [2]
    uint64_t* tmem_empty = tmem_full + 2;                              // epilogue -> MMA   [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.tmem_slot);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < raw_stages; ++s) { tc_mbar_init(&full[s], 1); tc_mbar_init(&raw_free[s], 4); }
        for (int s = 0; s < kTcStages; ++s) { tc_mbar_init(&ready[s], 4); tc_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { tc_mbar_init(&tmem_full[b], 1); tc_mbar_init(&tmem_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                                   // TMEM allocation (this warp also frees it)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(kTcTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Y -> shared memory, split into hi / lo, K-major no-swizzle core matrices: element (k = i, n = r) of chunk c = i / 8 at
    //   c * n_cols * 8  +  ((i % 8) / 4) * (n_cols / 8) * 32  +  (r / 8) * 32  +  (r % 8) * 4  +  (i % 4)        [floats]
    for (int idx = tid; idx < kchunks_pad * 8 * n_cols; idx += kTcThreads) {
        const int i = idx / n_cols, r = idx - i * n_cols;
        const float y = (i < m && r < nvec) ? Y[(int64_t)i * ldy + v0 + r] : 0.0f;
        const float hi = tf32_hi(y);
        const int off = (i >> 3) * n_cols * 8 + (((i & 7) >> 2) * (n_cols >> 3) + (r >> 3)) * 32 + (r & 7) * 4 + (i & 3);
        b_hi[off] = hi;
        b_lo[off] = tf32_hi(y - hi);
    }
    tc_fence_proxy_async();                                            // the tensor core reads Y through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int x0 = (int)(tile * kTcTileCols);
                for (int ks = 0; ks < nstages_per_tile; ++ks, ++it) {
                    const int s = it % raw_stages;
                    tc_mbar_wait(&raw_free[s], ((it / raw_stages) & 1u) ^ 1u);
                    tc_mbar_expect_tx(&full[s], kTcStageBytes);
                    tc_tma_box(raw + (size_t)s * kTcStageBytes, &tmap, x0, ks * kTcStageRows, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = tc_idesc(n_cols);
        const uint32_t b_lbo = (uint32_t)(n_cols >> 3) * 128u, b_chunk = (uint32_t)n_cols * 32u;
        uint32_t it = 0, tcount = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const uint32_t acc = tcount & 1u;
            tc_mbar_wait(&tmem_empty[acc], ((tcount >> 1) & 1u) ^ 1u);     // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_addr = tmem_base + acc * 128u;
            for (int ks = 0; ks < nstages_per_tile; ++ks, ++it) {
                const int s = it % kTcStages;
                tc_mbar_wait(&ready[s], (it / kTcStages) & 1u);
                tc_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const uint32_t kc = (uint32_t)(ks * 2 + a);
                        const uint64_t db_hi = tc_desc(tc_smem_u32(b_hi) + kc * b_chunk, b_lbo, 128, 0);
                        const uint64_t db_lo = tc_desc(tc_smem_u32(b_lo) + kc * b_chunk, b_lbo, 128, 0);
                        if constexpr (A_TMEM) {
                            // K atom a of the stage: 8 columns of the (hi | lo) operand stage in tensor memory
                            const uint32_t ah = tmem_base + kTcTmemA + (uint32_t)s * 32u + (uint32_t)a * 8u, al = ah + 16u;
                            tc_mma_tf32_ts(d_addr, ah, db_hi, idesc, (ks | a) ? 1u : 0u);
                            tc_mma_tf32_ts(d_addr, al, db_hi, idesc, 1u);
                            tc_mma_tf32_ts(d_addr, ah, db_lo, idesc, 1u);
                        } else {
                            // K atom a = k-quads 2a, 2a+1 of the stage (2 KB each): LBO = 2 KB between the quads, SBO = 128 B
                            // between the 8-column groups of core-matrix rows
                            const uint64_t da_hi = tc_desc(tc_smem_u32(a_hi + (size_t)s * kTcStageBytes + a * 4096), 2048, 128, 0);
                            const uint64_t da_lo = tc_desc(tc_smem_u32(a_lo + (size_t)s * kTcStageBytes + a * 4096), 2048, 128, 0);
                            tc_mma_tf32(d_addr, da_hi, db_hi, idesc, (ks | a) ? 1u : 0u);
                            tc_mma_tf32(d_addr, da_lo, db_hi, idesc, 1u);
                            tc_mma_tf32(d_addr, da_hi, db_lo, idesc, 1u);
                        }
                    }
                    tc_commit(&empty[s]);                               // stage free once these MMAs have read it
                    if (ks == nstages_per_tile - 1) tc_commit(&tmem_full[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ===== transpose + split warps: raw [16 rows][128 columns] -> q_hi, q_lo rows of the A operand in TENSOR MEMORY =====
        // Thread = column x of the tile = TMEM lane (a warp may touch the lane quarter (warp % 4) only): it reads its 16 values
        // of the stage from shared memory (conflict-free), splits them and stores both halves with one tcgen05.st each --
        // the transpose costs nothing, and the MMA reads A from tensor memory instead of shared memory.
        const int x = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTcTmemA;
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int ks = 0; ks < nstages_per_tile; ++ks, ++it) {
                const int rs = it % raw_stages, s = it % kTcStages;
                tc_mbar_wait(&full[rs], (it / raw_stages) & 1u);       // the raw tile has landed
                const float* src = reinterpret_cast<const float*>(raw + (size_t)rs * kTcStageBytes) + x;
                float h[kTcStageRows], l[kTcStageRows];
#pragma unroll
                for (int i = 0; i < kTcStageRows; ++i) {
                    const float q = src[i * kTcTileCols];
                    h[i] = tf32_hi(q);
                    l[i] = tf32_hi(q - h[i]);
                }
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&raw_free[rs]);          // registers hold it: the producer may refill the slot
                tc_mbar_wait(&empty[s], ((it / kTcStages) & 1u) ^ 1u); // the MMAs that read this operand stage are done
                if constexpr (A_TMEM) {
                    tc_fence_after();
                    tc_st16(lane_addr + (uint32_t)s * 32u, h);
                    tc_st16(lane_addr + (uint32_t)s * 32u + 16u, l);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                } else {
                    // K-major no-swizzle core matrices: (x, i) at (i / 4) * 2 KB + (x / 8) * 128 + (x % 8) * 16 + (i % 4) * 4
                    unsigned char* hi = a_hi + (size_t)s * kTcStageBytes + (x >> 3) * 128 + (x & 7) * 16;
                    unsigned char* lo = a_lo + (size_t)s * kTcStageBytes + (x >> 3) * 128 + (x & 7) * 16;
#pragma unroll
                    for (int kq = 0; kq < kTcStageRows / 4; ++kq) {
                        *reinterpret_cast<float4*>(hi + kq * 2048) = make_float4(h[4 * kq], h[4 * kq + 1], h[4 * kq + 2], h[4 * kq + 3]);
                        *reinterpret_cast<float4*>(lo + kq * 2048) = make_float4(l[4 * kq], l[4 * kq + 1], l[4 * kq + 2], l[4 * kq + 3]);
                    }
                    tc_fence_proxy_async();                            // generic-proxy writes -> visible to the tensor core
                }
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(&ready[s]);
            }
        }
    } else {
        // ===== epilogue warps: TMEM -> registers -> global =====
        const int q = warp & 3;                                        // TMEM lane quarter this warp may read
        uint32_t tcount = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const uint32_t acc = tcount & 1u;
            tc_mbar_wait(&tmem_full[acc], (tcount >> 1) & 1u);
            tc_fence_after();
            const int64_t x = tile * kTcTileCols + q * 32 + lane;
            const uint32_t taddr = tmem_base + acc * 128u + ((uint32_t)(q * 32) << 16);
            float* dst = out + (int64_t)v0 * ldo + x;
            // every load of the tile is issued before ONE wait (tcgen05.ld shares the tensor core's in-order queue with the MMAs of
            // the next tile: a wait per chunk would queue behind them again and again); stores of full chunks carry no predicate
            uint32_t v[kTcMaxN / 16][16];
#pragma unroll
            for (int k = 0; k < kTcMaxN / 16; ++k)
                if (k * 16 < n_cols) tc_ld16_nowait(taddr + k * 16, v[k]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&tmem_empty[acc]);          // the accumulator is in registers: the MMA warp may reuse it
#pragma unroll
            for (int k = 0; k < kTcMaxN / 16; ++k) {
                if (k * 16 + 16 <= nvec) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) dst[(int64_t)(k * 16 + j) * ldo] = __uint_as_float(v[k][j]);
                } else if (k * 16 < nvec) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (k * 16 + j < nvec) dst[(int64_t)(k * 16 + j) * ldo] = __uint_as_float(v[k][j]);
                }
            }
        }
    }
    // ===== teardown =====
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTcTmemCols) : "memory");
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda).
typedef CUresult (*TcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TcEncodeTiledFn tc_encode_tiled() {
    static TcEncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<TcEncodeTiledFn>(p);
    }();
    return fn;
}

// Columns [0, n_main) of out[v0 .. v0+nvec) on the tensor cores; returns n_main (a multiple of 128; 0 = not applicable).
int ritz_vectors_tc(const float* Q, int64_t ldq, int m, const float* Y, int ldy, int v0, int nvec, float* out, int64_t ldo,
                    int64_t n, cudaStream_t stream, int64_t* n_main_out) {
    *n_main_out = 0;
    static const bool enabled = [] { const char* e = getenv("HLV_RITZ_TC"); return !(e && e[0] == '0'); }();
    if (!enabled) return HLV_OK;
    const int64_t ntiles = n / kTcTileCols;
    if (m > kTcMaxM || nvec > kTcMaxN || nvec < 1 || ntiles < 1 || n > (int64_t)INT32_MAX - 4096) return HLV_OK;
    TcEncodeTiledFn enc = tc_encode_tiled();
    if (enc == nullptr) return HLV_OK;
    const int n_cols = nvec <= 16 ? 16 : (nvec + 15) / 16 * 16;
    const int kchunks_pad = 2 * (((m + 7) / 8 + 1) / 2);
    // HLV_RITZ_A=smem keeps the split operand in shared memory (tcgen05.mma SS form) instead of tensor memory (TS form, default);
    // both are validated by the tests, measured within 3% of each other (24.0 / 24.6 ms for all 100 vectors at GPT-2 size)
    static const bool a_tmem = [] { const char* e = getenv("HLV_RITZ_A"); return !(e && e[0] == 's'); }();
    int raw_stages = kTcMaxRawStages;                                                // as deep as the 227 KB allow
    while (raw_stages > 2 && ritz_tc_layout(kchunks_pad, n_cols, raw_stages, a_tmem).total + 1024 > 227 * 1024) --raw_stages;
    const size_t smem = ritz_tc_layout(kchunks_pad, n_cols, raw_stages, a_tmem).total + 1024;   // + slack for the alignment of the carve-up
    if (smem > 227 * 1024) return HLV_OK;
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)(ntiles * kTcTileCols), (cuuint64_t)m};
    const cuuint64_t gstride[1] = {(cuuint64_t)ldq * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kTcTileCols, (cuuint32_t)kTcStageRows};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Q), gdim, gstride, box, estride,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HLV_REQUIRE(r == CUDA_SUCCESS, HLV_ERR_ARG, "hlv_ritz_vectors_f32: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    int64_t grid = sm_count();
    if (grid > ntiles) grid = ntiles;
    cudaError_t e;
    if (a_tmem) {
        if ((e = ensure_dynamic_smem(reinterpret_cast<const void*>(ritz_vectors_tc_kernel<true>), smem)) != cudaSuccess)
            return cuda_fail(e, "cudaFuncSetAttribute(ritz_vectors_tc)");
        ritz_vectors_tc_kernel<true><<<(int)grid, kTcThreads, smem, stream>>>(map, m, Y, ldy, v0, nvec, n_cols, out, ldo, ntiles, raw_stages);
    } else {
        if ((e = ensure_dynamic_smem(reinterpret_cast<const void*>(ritz_vectors_tc_kernel<false>), smem)) != cudaSuccess)
            return cuda_fail(e, "cudaFuncSetAttribute(ritz_vectors_tc)");
        ritz_vectors_tc_kernel<false><<<(int)grid, kTcThreads, smem, stream>>>(map, m, Y, ldy, v0, nvec, n_cols, out, ldo, ntiles, raw_stages);
    }
    HLV_LAUNCH_CHECK("hlv_ritz_vectors_f32 (tensor-core pass)");
    *n_main_out = ntiles * kTcTileCols;
    return HLV_OK;
}

}  // namespace hlv
