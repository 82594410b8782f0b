// ws_mailbox.cuh — the small exchanges of a sharded resampling step, done by the kernels themselves over NVLink.
//
// A sharded step needs four tiny all-to-all exchanges (the ranks' (m, S, Q) triples, their fixed-point masses, the
// slot bounds + plane addresses of the exchange plan, and a barrier behind the pushed offspring).  As NCCL
// collectives each costs a launch and a latency-bound kernel of its own; here every rank owns a MAILBOX in device
// memory that all other ranks map into their address space (cudaIpc, once, at ws_create_sharded), and the kernel that
// produces a message stores it straight into the peers' mailboxes and spins on its own until everybody's has arrived —
// so "reduce, exchange, combine" and "offsets, exchange, bounds, exchange" are ONE kernel each.
//
// Protocol (the shape of NCCL's LL protocol): a message of W 64-bit words travels as 2 W eight-byte stores, each
// carrying 32 bits of payload and the 32-bit sequence number of the exchange.  An aligned 8-byte store is single-copy
// atomic, so a receiver that reads a word with the expected sequence number has its payload — no fence, no separate
// flag, one NVLink traversal.  Mailboxes are double-buffered on the parity of the sequence number: every exchange is
// all-to-all, so a rank can start exchange k + 2 only after it has received every rank's message k + 1, which a rank
// sends only after it has finished reading exchange k.  All ranks run the same sequence of exchanges (SPMD); a gated
// exchange that does not fire is skipped by everybody (the decision is bit-identical on all ranks) and its sequence
// numbers are simply never seen.
//
// A rank that waits longer than `timeout_ns` (a peer died, the ranks left lock-step) raises *err (mapped host memory),
// stops waiting and lets the host fail the call — a spin never outlives the job.
#pragma once
#include <stdint.h>

#define WS_MBOX_MAX_RANKS 16
#define WS_MBOX_THREADS 256

struct WsMailbox {
    unsigned long long* box[WS_MBOX_MAX_RANKS];  // rank q's mailbox in THIS process's address space (box[rank]: my own)
    int32_t rank, nranks;
    uint32_t seq;                                // sequence number of the kernel's first exchange (> 0)
    int32_t cap;                                 // 8-byte words per (parity, source) region
    unsigned int* err;                           // != 0: an exchange timed out
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long ws_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void ws_st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ws_ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// All threads of the (single) CTA call.  `mine`: n_words 64-bit words of this rank (global or shared memory, already
// visible to the whole CTA); out[q * n_words + w] (global or shared) receives rank q's words, q in rank order, for all
// R ranks including this one.  Ends with a CTA barrier: afterwards every thread may read `out`.
__device__ __forceinline__ void ws_mbox_allgather(const WsMailbox& M, const uint32_t seq, const unsigned long long* mine, const int n_words,
                                                  unsigned long long* out) {
    const int R = M.nranks, ll = 2 * n_words, total = ll * R;
    const size_t region = (size_t)(seq & 1u) * (size_t)R;
    const unsigned long long tag = (unsigned long long)seq << 32;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int q = i / ll, w = i - q * ll;
        const unsigned long long half = (mine[w >> 1] >> (32 * (w & 1))) & 0xFFFFFFFFull;
        ws_st_sys_u64(M.box[q] + (region + (size_t)M.rank) * (size_t)M.cap + w, tag | half);
    }
    const unsigned long long* const self = M.box[M.rank];
    uint32_t* const out32 = reinterpret_cast<uint32_t*>(out);
    unsigned long long t0 = 0ull;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int q = i / ll, w = i - q * ll;
        const unsigned long long* p = self + (region + (size_t)q) * (size_t)M.cap + w;
        unsigned long long v = ws_ld_sys_u64(p);
        unsigned int spins = 0u;
        while ((uint32_t)(v >> 32) != seq) {
            if ((++spins & 1023u) == 0u) {
                const unsigned long long now = ws_globaltimer();
                if (t0 == 0ull) t0 = now;
                if (now - t0 > M.timeout_ns || *reinterpret_cast<volatile unsigned int*>(M.err) != 0u) {
                    *reinterpret_cast<volatile unsigned int*>(M.err) = 1u;
                    break;
                }
            }
            v = ws_ld_sys_u64(p);
        }
        out32[(size_t)q * ll + w] = (uint32_t)v;
    }
    __syncthreads();
}

// Barrier over the ranks behind peer stores of EARLIER kernels of this stream (the pushed offspring): the fence makes
// them visible system-wide before this rank's word can be seen.
__device__ __forceinline__ void ws_mbox_barrier(const WsMailbox& M, const uint32_t seq, unsigned long long* scratch /* [nranks], shared */) {
    __shared__ unsigned long long one;
    if (threadIdx.x == 0) one = 1ull;
    __threadfence_system();
    __syncthreads();
    ws_mbox_allgather(M, seq, &one, 1, scratch);
    __threadfence_system();
}
