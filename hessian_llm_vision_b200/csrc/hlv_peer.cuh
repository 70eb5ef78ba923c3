// Scalar / flag exchange between ranks over NVLink peer memory, used from INSIDE the recurrence kernels.
//
// Every rank owns one "exchange area" (hlv_peer_xchg_bytes() bytes of peer-mapped memory); rank r knows the
// address of every rank's area as mapped into ITS address space (hlv_peer_ctx::xchg).  Layout of one area:
//
//   sent [kChannels]                     u64   local bookkeeping: number of pushes this rank has made per channel
//   error                                u32   set when a wait ran into its time limit (results are then invalid)
//   flags[kChannels][HLV_MAX_PEERS]      u64   flags[c][src] = epoch of src's latest push on channel c   (written by src)
//   slots[kChannels][2][HLV_MAX_PEERS][kSlotDoubles]  f64  payload of that push, double-buffered by epoch parity
//
// A "push" is done by ONE CTA (the one that finalises a kernel's cross-CTA reduction): it stores its k partial values
// into slot [c][epoch&1][my rank] of EVERY rank's area with plain peer stores, fences at system scope and then raises
// flags[c][my rank] = epoch on every rank (st.release.sys).  A "pull" is done by EVERY CTA of the consuming kernel:
// wait until all `world` flags of the channel have reached this rank's own push count, then add the `world` slots in
// rank order in float64 -- every rank adds the same numbers in the same order, so the replicated scalars (alpha,
// beta, Gram-Schmidt coefficients) stay bit-identical across ranks, and independent of arrival order.
// This replaces a k-float all-reduce LAUNCH per reduction by a few hundred bytes of peer stores in the producer's
// epilogue and a few L2 reads in the consumer's prologue.
//
// Why the double buffering is enough: on every rank a channel's pushes and pulls alternate in stream order (push e,
// pull e, push e+1, ...).  Rank A can push epoch e+2 only after it pulled e+1, i.e. after rank B pushed e+1, i.e.
// after B finished the kernel that pulled e -- so nobody still reads parity (e & 1) when it is overwritten.
//
// Each rank runs on its own GPU: a spinning consumer never occupies the SMs its producer needs.  (Ranks must NOT
// share a GPU; tests emulate several ranks on one device only with every push issued before the matching pull.)
#pragma once

#include "hlv_common.cuh"

namespace hlv {

constexpr int kChannels = HLV_PEER_CHANNELS;
constexpr int kSlotDoubles = HLV_MAX_ROWS + 8;
constexpr size_t kXchgSentOff = 0;                                            // u64[kChannels]
constexpr size_t kXchgErrorOff = 128;                                         // u32
constexpr size_t kXchgFlagsOff = 256;                                         // u64[kChannels][HLV_MAX_PEERS]
constexpr size_t kXchgSlotsOff = kXchgFlagsOff + sizeof(unsigned long long) * kChannels * HLV_MAX_PEERS;
constexpr size_t kXchgBytes = kXchgSlotsOff + sizeof(double) * kChannels * 2 * HLV_MAX_PEERS * kSlotDoubles;
static_assert(kXchgSlotsOff % 16 == 0, "slot alignment");

// Kernel-side view of hlv_peer_ctx (passed by value).  world <= 1: single rank, every peer operation degenerates to
// "use the local value".
struct PeerView {
    int world, rank;
    unsigned long long timeout_ns;
    char* xchg[HLV_MAX_PEERS];
};

inline PeerView make_peer_view(const hlv_peer_ctx* h) {
    PeerView v{};
    v.world = 1;
    if (h != nullptr && h->world > 1) {
        v.world = h->world;
        v.rank = h->rank;
        v.timeout_ns = (unsigned long long)(h->spin_timeout_ms ? h->spin_timeout_ms : 20000u) * 1000000ull;
        for (int p = 0; p < h->world && p < HLV_MAX_PEERS; ++p) v.xchg[p] = static_cast<char*>(h->xchg[p]);
    }
    return v;
}
inline int check_peer_ctx(const hlv_peer_ctx* h, const char* name) {
    if (h == nullptr) return HLV_OK;
    HLV_REQUIRE(h->world >= 1 && h->world <= HLV_MAX_PEERS && h->rank >= 0 && h->rank < h->world, HLV_ERR_ARG,
                "%s: peer context world=%d rank=%d out of range (max %d ranks)", name, h->world, h->rank, HLV_MAX_PEERS);
    for (int p = 0; p < h->world; ++p)
        HLV_REQUIRE(h->world == 1 || (h->xchg[p] != nullptr && aligned16(h->xchg[p])), HLV_ERR_ARG,
                    "%s: peer context has no (aligned) exchange area for rank %d", name, p);
    return HLV_OK;
}

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long* xchg_sent(char* base, int c) {
    return reinterpret_cast<unsigned long long*>(base + kXchgSentOff) + c;
}
__device__ __forceinline__ unsigned* xchg_error(char* base) { return reinterpret_cast<unsigned*>(base + kXchgErrorOff); }
__device__ __forceinline__ unsigned long long* xchg_flag(char* base, int c, int src) {
    return reinterpret_cast<unsigned long long*>(base + kXchgFlagsOff) + (size_t)c * HLV_MAX_PEERS + src;
}
__device__ __forceinline__ double* xchg_slot(char* base, int c, int parity, int src) {
    return reinterpret_cast<double*>(base + kXchgSlotsOff) + (((size_t)c * 2 + parity) * HLV_MAX_PEERS + src) * kSlotDoubles;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Epoch this rank's NEXT pull on channel c has to wait for = the number of pushes it has made itself (its own push
// of that epoch was issued by an earlier kernel of the same stream, or earlier in this kernel by the same CTA).
__device__ __forceinline__ unsigned long long peer_epoch(const PeerView& pv, int c) {
    return ld_relaxed_u64(xchg_sent(pv.xchg[pv.rank], c));
}

// ONE CTA, all of its threads.  vals[0..k) must be visible to the CTA (caller synchronised).  k == 0: flag only.
__device__ __forceinline__ void peer_push(const PeerView& pv, int c, const double* vals, int k) {
    if (pv.world <= 1) return;
    char* mine = pv.xchg[pv.rank];
    const unsigned long long e = ld_relaxed_u64(xchg_sent(mine, c)) + 1ull;
    const int parity = (int)(e & 1ull);
    for (int idx = threadIdx.x; idx < pv.world * k; idx += blockDim.x) {
        const int p = idx / k, i = idx - p * k;
        xchg_slot(pv.xchg[p], c, parity, pv.rank)[i] = vals[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pv.world) st_release_sys(xchg_flag(pv.xchg[threadIdx.x], c, pv.rank), e);
    if (threadIdx.x == 0) *xchg_sent(mine, c) = e;
}

// EVERY CTA, all of its threads: wait for all ranks' push of the current epoch on channel c.  Returns the epoch.
__device__ __forceinline__ unsigned long long peer_wait_all(const PeerView& pv, int c) {
    char* mine = pv.xchg[pv.rank];
    const unsigned long long e = peer_epoch(pv, c);
    if ((int)threadIdx.x < pv.world) {
        const unsigned long long* f = xchg_flag(mine, c, threadIdx.x);
        if (ld_acquire_sys(f) < e) {
            const unsigned long long t0 = global_timer_ns();
            unsigned spins = 0;
            while (ld_acquire_sys(f) < e) {
                __nanosleep(64);
                if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > pv.timeout_ns) {
                    atomicExch(xchg_error(mine), 1u + (unsigned)c);      // give up: the host reports it, results are invalid
                    break;
                }
            }
        }
    }
    __syncthreads();
    return e;
}

// EVERY CTA: emit(i, total_i) for i < k, total_i = sum over ranks (in rank order, float64) of the values pushed on
// channel c.  Ends with a __syncthreads().
template <typename Emit>
__device__ __forceinline__ void peer_pull(const PeerView& pv, int c, int k, Emit emit) {
    const unsigned long long e = peer_wait_all(pv, c);
    char* mine = pv.xchg[pv.rank];
    const int parity = (int)(e & 1ull);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        double s = 0.0;
        for (int p = 0; p < pv.world; ++p) s += ld_relaxed_sys_f64(xchg_slot(mine, c, parity, p) + i);
        emit(i, s);
    }
    __syncthreads();
}
// One scalar for every thread of the CTA (alpha, |w|^2).
__device__ __forceinline__ double peer_pull_scalar(const PeerView& pv, int c) {
    const unsigned long long e = peer_wait_all(pv, c);
    char* mine = pv.xchg[pv.rank];
    const int parity = (int)(e & 1ull);
    double s = 0.0;
    for (int p = 0; p < pv.world; ++p) s += ld_relaxed_sys_f64(xchg_slot(mine, c, parity, p));
    return s;
}

// Cross-CTA final stage with an optional push.  Same contract as finalize_rows (hlv_common.cuh); when `pv.world > 1`
// the finalising CTA also pushes the `rows` totals on channel `push_channel` (out[] still receives the LOCAL totals).
__device__ __forceinline__ void finalize_rows_push(const double* partials, unsigned* counter, int rows, double* out,
                                                   const PeerView& pv, int push_channel) {
    __shared__ unsigned s_ticket_p;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket_p = atomicInc(counter, gridDim.x - 1);
    __syncthreads();
    if (s_ticket_p != gridDim.x - 1) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int nblk = gridDim.x;
    for (int r = warp; r < rows; r += nwarps) {
        const double* p = partials + (size_t)r * kMaxCtas;
        double s = 0.0;
        for (int b = lane; b < nblk; b += 32) s += __ldcg(p + b);
        s = warp_sum(s);
        if (lane == 0) out[r] = s;
    }
    if (pv.world > 1 && push_channel >= 0) {
        __syncthreads();
        peer_push(pv, push_channel, out, rows);
    }
}
#endif  // __CUDACC__

}  // namespace hlv
