// ws_kernels_move.cu — score-tape fold and fused Metropolis-Hastings move kernels.
//
//   ws_score_kernel         score_logpdf(state, targets, depth)            src/types.jl:183-206
//   ws_move_kernel          Move.apply!: propose / score old / score new / accept-or-restore in
//                           one pass                                       src/transformers.jl:588-623
//                           RW / autoRW proposals and bound transforms      src/move_kernels.jl:37-85,161-253
//   ws_move_moments_kernel  weighted mean / covariance of the (unconstrained) targets for autoRW
//                                                                          src/move_kernels.jl:144-151
//   ws_unique_count_kernel  exact distinct-value count for marginal_diversity
//                                                                          src/transformers.jl:560-565
//
// The fold is FP64-ALU bound for long tapes (one thread per particle walks k tape entries; the
// particle's planes sit in the shared-memory register file, the micro-ops are fetched with
// warp-uniform loads that hit L1/L2), so it is reported against FP64 issue rate, not HBM.
#include "ws_move.h"

__device__ __forceinline__ WsOp ws_load_op(const WsOp* __restrict__ ops, int pc) {
    // two warp-uniform 16-byte loads
    const uint4* p = reinterpret_cast<const uint4*>(ops + pc);
    uint4 a = __ldg(p), b = __ldg(p + 1);
    WsOp o;
    o.w0 = a.x;
    o.w1 = a.y;
    o.k0 = __hiloint2double((int)a.w, (int)a.z);
    o.k1 = __hiloint2double((int)b.y, (int)b.x);
    o.k2 = __hiloint2double((int)b.w, (int)b.z);
    return o;
}

// P particles per thread: particle j of a thread is element tile*P*BLOCK + j*BLOCK + tid (coalesced), register r of
// particle j lives at R[(r*P + j)*BLOCK].  One decoded tape entry is applied to P particles, which is what makes a
// long fold cheap: the interpreter's decode (~30 instructions) dwarfs the 3-4 FP64 operations of a term.
template <int P>
__device__ __forceinline__ void ws_score_load_planes(const WsScoreParams& S, double* R, const int64_t (&idx)[P]) {
    for (int k = 0; k < S.n_loads; ++k) {
        const double* __restrict__ ptr = S.load_ptr[k];
        double t[P];
#pragma unroll
        for (int j = 0; j < P; ++j) t[j] = __ldg(ptr + idx[j]);
        double* dst = R + (int)S.load_reg[k] * (P * WS_MOVE_BLOCK);
#pragma unroll
        for (int j = 0; j < P; ++j) dst[j * WS_MOVE_BLOCK] = t[j];
    }
}

// The fold: every thread of the CTA walks the same tape, so the CTA stages it through shared memory in chunks of
// WS_FOLD_CHUNK entries, each entry unpacked once (ws_vm.cuh: WsDop) by one thread instead of by every thread for
// every tile.  Must be called by all threads of the CTA (two barriers per chunk).
#define WS_FOLD_CHUNK WS_MOVE_BLOCK
template <int P>
__device__ __forceinline__ void ws_score_fold(const WsScoreParams& S, double* R, const uint64_t (&pid)[P], double (&acc)[P],
                                              WsDop* dops /* [WS_FOLD_CHUNK], shared */) {
#pragma unroll
    for (int j = 0; j < P; ++j) acc[j] = 0.0;
    WsRng none;
    none.seed = 0;
    none.replay_n = nullptr;
    none.replay_u = nullptr;
    none.replay_e = nullptr;
    none.replay_v = nullptr;
    __shared__ unsigned int cont_mask[WS_FOLD_CHUNK / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < S.n_ops; base += WS_FOLD_CHUNK) {
        const int cnt = min(WS_FOLD_CHUNK, S.n_ops - base);
        __syncthreads();  // the previous chunk (or fold) has been consumed by every warp
        // thread t unpacks entry t; `cont`: a squared-residual entry over the same registers as its predecessor
        bool cont = false, runnable = false;
        if ((int)threadIdx.x < cnt) {
            const WsOp o = ws_load_op(S.ops, base + (int)threadIdx.x);
            WsDop dd = ws_decode_op<WS_MOVE_BLOCK, P>(o, S.reg_map);
            runnable = dd.op == WS_OP_ACC_SQLIN2 || dd.op == WS_OP_ACC_SQLIN2_S;
            if (runnable && threadIdx.x > 0) {
                const uint2 pw = __ldg(reinterpret_cast<const uint2*>(S.ops + base + (int)threadIdx.x - 1));
                WsOp prev;
                prev.w0 = pw.x;
                prev.w1 = pw.y;
                cont = ws_run_continues(prev, o);
            }
            dd.op |= 1u << 8;
            dops[threadIdx.x] = dd;
        }
        const unsigned int cm = __ballot_sync(0xffffffffu, cont);
        if (lane == 0) cont_mask[warp] = cm;
        __syncthreads();
        if (runnable && !cont) {  // head of a run: its length = 1 + the `cont` bits that follow
            int len = 1, q = (int)threadIdx.x + 1;
            while (q < WS_FOLD_CHUNK) {
                const unsigned int inv = ~(cont_mask[q >> 5] >> (q & 31));
                const int avail = 32 - (q & 31);
                const int ones = inv ? __ffs((int)inv) - 1 : 32;
                if (ones >= avail) {
                    len += avail;
                    q += avail;
                } else {
                    len += ones;
                    break;
                }
            }
            if (len > 1) dops[threadIdx.x].op = (dops[threadIdx.x].op & 0xFFu) | ((unsigned int)len << 8);
        }
        __syncthreads();
        // the next entry is fetched while the current one executes (the fetch -> dispatch chain is otherwise
        // exposed: few warps per scheduler)
        WsDop cur = dops[0];
        for (int k = 0; k < cnt;) {
            const int len = (int)(cur.op >> 8);
            const int kn = k + len;
            const WsDop nxt = dops[kn < cnt ? kn : k];
            if (len > 1) ws_vm_exec_sqlin2_run<WS_MOVE_BLOCK, P>(dops + k, len, R, acc);
            else ws_vm_exec_d<WS_MOVE_BLOCK, P>(cur, R, acc, none, pid);
            cur = nxt;
            k = kn;
        }
    }
}

template <int P>
__device__ __forceinline__ void ws_tile_indices(int64_t tile, int64_t n, int64_t offset, int64_t (&idx)[P], bool (&live)[P],
                                                uint64_t (&pid)[P]) {
#pragma unroll
    for (int j = 0; j < P; ++j) {
        const int64_t i = tile * (P * WS_MOVE_BLOCK) + (int64_t)j * WS_MOVE_BLOCK + threadIdx.x;
        live[j] = i < n;
        idx[j] = live[j] ? i : n - 1;
        pid[j] = (uint64_t)(offset + idx[j]);
    }
}

template <int P>
__global__ void __launch_bounds__(WS_MOVE_BLOCK) ws_score_kernel(const __grid_constant__ WsScoreParams S) {
    extern __shared__ double ws_score_smem[];
    __shared__ WsDop dops[WS_FOLD_CHUNK];
    double* R = ws_score_smem + threadIdx.x;
    const int64_t n_tiles = (S.n + P * WS_MOVE_BLOCK - 1) / (P * WS_MOVE_BLOCK);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t idx[P];
        bool live[P];
        uint64_t pid[P];
        ws_tile_indices<P>(tile, S.n, S.particle_offset, idx, live, pid);
        ws_score_load_planes<P>(S, R, idx);
        double acc[P];
        ws_score_fold<P>(S, R, pid, acc, dops);
#pragma unroll
        for (int j = 0; j < P; ++j)
            if (live[j]) S.score_out[idx[j]] = acc[j] + S.konst;
    }
}

// Rows [0, n_regs) of the register file are the score program's registers, rows [n_regs, n_regs + d) hold the
// proposed target values of the thread's P particles.
#ifndef WS_MOVE_MINB
#define WS_MOVE_MINB 4   // <= 128 registers: 16 warps per SM (measured against 1 and 3, scripts/ab_move.sh)
#endif
template <int P>
__global__ void __launch_bounds__(WS_MOVE_BLOCK, WS_MOVE_MINB) ws_move_kernel(const __grid_constant__ WsMoveParams M) {
    extern __shared__ double ws_score_smem[];
    __shared__ WsDop dops[WS_FOLD_CHUNK];
    double* R = ws_score_smem + threadIdx.x;
    constexpr int RS = P * WS_MOVE_BLOCK;
    const WsScoreParams& S = M.score;
    const int d = M.d;
    double* const Xn = R + S.n_regs * RS;
    unsigned long long accepted = 0ull;
    const int64_t n_tiles = (S.n + RS - 1) / RS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t idx[P];
        bool live[P];
        uint64_t pid[P];
        ws_tile_indices<P>(tile, S.n, S.particle_offset, idx, live, pid);
        ws_score_load_planes<P>(S, R, idx);

        // ---- current values, unconstrained coordinates, proposal (one particle at a time) -----------
        double lpr[P];
#pragma unroll 1
        for (int j = 0; j < P; ++j) {
            const int64_t i = idx[j];
            const uint64_t particle = pid[j];
            double z_old[WS_MOVE_MAX_D], xi[WS_MOVE_MAX_D];
#pragma unroll
            for (int t = 0; t < WS_MOVE_MAX_D; ++t)
                if (t < d) z_old[t] = ws_to_unconstrained(M.target_ptr[t][i], M.lo[t], M.hi[t], M.bound_kind[t]);
            if (M.rng.replay_n != nullptr) {
#pragma unroll
                for (int t = 0; t < WS_MOVE_MAX_D; ++t) {
                    if (t < d) {
                        const int64_t ri = M.normals_target_major ? (M.replay_n_base + (int64_t)t * M.n_global + (int64_t)particle)
                                                                  : (M.replay_n_base + (int64_t)particle * d + t);
                        xi[t] = M.rng.replay_n[ri];
                    }
                }
            } else {
#pragma unroll
                for (int t = 0; t < WS_MOVE_MAX_D; t += 2) {
                    if (t < d) {
                        double a, b;
                        ws_randn2(particle, M.stream_normals + (uint64_t)(t >> 1), M.rng.seed, a, b);
                        xi[t] = a;
                        if (t + 1 < WS_MOVE_MAX_D) xi[t + 1] = b;
                    }
                }
            }
            double l = 0.0;
#pragma unroll
            for (int t = 0; t < WS_MOVE_MAX_D; ++t) {
                if (t < d) {
                    // sequential, un-contracted sum so that a replayed proposal is bit-identical to
                    // the reference's  z .+ (L * xi)  (no FMA in Julia)
                    double delta = __dmul_rn(M.L[t * d], xi[0]);
#pragma unroll
                    for (int k = 1; k < WS_MOVE_MAX_D; ++k)
                        if (k <= t && k < d) delta = __dadd_rn(delta, __dmul_rn(M.L[t * d + k], xi[k]));
                    const double zn = __dadd_rn(z_old[t], delta);
                    Xn[t * RS + j * WS_MOVE_BLOCK] = ws_from_unconstrained(zn, M.lo[t], M.hi[t], M.bound_kind[t]);
                    l += ws_log_abs_jacobian(zn, M.lo[t], M.hi[t], M.bound_kind[t]) -
                         ws_log_abs_jacobian(z_old[t], M.lo[t], M.hi[t], M.bound_kind[t]);
                }
            }
            lpr[j] = l;
        }

        // ---- trace density at the old and at the proposed values ------------------------------------
        double s_old[P], s_new[P];
        ws_score_fold<P>(S, R, pid, s_old, dops);
        for (int t = 0; t < d; ++t) {
            if (M.target_reg[t] != 0xFF) {
                double* dst = R + (int)M.target_reg[t] * RS;
#pragma unroll
                for (int j = 0; j < P; ++j) dst[j * WS_MOVE_BLOCK] = Xn[t * RS + j * WS_MOVE_BLOCK];
            }
        }
        ws_score_fold<P>(S, R, pid, s_new, dops);

        // ---- accept / reject (NaN ratio rejects: !(log u < ...)) ---------------------------------------
#pragma unroll
        for (int j = 0; j < P; ++j) {
            if (!live[j]) continue;
            double u;
            if (M.rng.replay_u != nullptr) {
                u = M.rng.replay_u[M.replay_u_base + (int64_t)pid[j]];
            } else {
                ws_u32x4 r = ws_philox4x32_10(pid[j], M.stream_uniform, M.rng.seed);
                u = ws_u01(r.x, r.y);
            }
            if (log(u) < lpr[j] + s_new[j] - s_old[j]) {
                for (int t = 0; t < d; ++t) M.target_ptr[t][idx[j]] = Xn[t * RS + j * WS_MOVE_BLOCK];
                ++accepted;
            }
        }
    }
    // one atomic per warp
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) accepted += __shfl_down_sync(0xffffffffu, accepted, dlt);
    if ((threadIdx.x & 31) == 0 && accepted != 0ull) atomicAdd(M.n_accept, accepted);
}

// ------------------------------------------------------------------------------------------
// Wide tapes (more distinct planes than one register file holds): the move is split into
//   ws_move_propose_kernel   x' and the log proposal ratio into scratch
//   ws_move_delta_kernel     one launch per tape SEGMENT: delta += score(segment | x') - score(segment | x)
//   ws_move_accept_kernel    accept / restore
// The arithmetic per particle is the same as in ws_move_kernel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ws_move_propose_kernel(const __grid_constant__ WsMoveParams M, double* __restrict__ x_new,
                                                              double* __restrict__ lpr_out, double* __restrict__ delta) {
    const int d = M.d;
    const int64_t n = M.score.n;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const uint64_t particle = (uint64_t)(M.score.particle_offset + i);
        double z_old[WS_MOVE_MAX_D], xi[WS_MOVE_MAX_D];
#pragma unroll
        for (int t = 0; t < WS_MOVE_MAX_D; ++t)
            if (t < d) z_old[t] = ws_to_unconstrained(M.target_ptr[t][i], M.lo[t], M.hi[t], M.bound_kind[t]);
        if (M.rng.replay_n != nullptr) {
#pragma unroll
            for (int t = 0; t < WS_MOVE_MAX_D; ++t) {
                if (t < d) {
                    const int64_t idx = M.normals_target_major ? (M.replay_n_base + (int64_t)t * M.n_global + (int64_t)particle)
                                                               : (M.replay_n_base + (int64_t)particle * d + t);
                    xi[t] = M.rng.replay_n[idx];
                }
            }
        } else {
#pragma unroll
            for (int t = 0; t < WS_MOVE_MAX_D; t += 2) {
                if (t < d) {
                    double a, b;
                    ws_randn2(particle, M.stream_normals + (uint64_t)(t >> 1), M.rng.seed, a, b);
                    xi[t] = a;
                    if (t + 1 < WS_MOVE_MAX_D) xi[t + 1] = b;
                }
            }
        }
        double lpr = 0.0;
#pragma unroll
        for (int t = 0; t < WS_MOVE_MAX_D; ++t) {
            if (t < d) {
                double dl = __dmul_rn(M.L[t * d], xi[0]);
#pragma unroll
                for (int k = 1; k < WS_MOVE_MAX_D; ++k)
                    if (k <= t && k < d) dl = __dadd_rn(dl, __dmul_rn(M.L[t * d + k], xi[k]));
                const double zn = __dadd_rn(z_old[t], dl);
                x_new[(size_t)t * n + i] = ws_from_unconstrained(zn, M.lo[t], M.hi[t], M.bound_kind[t]);
                lpr += ws_log_abs_jacobian(zn, M.lo[t], M.hi[t], M.bound_kind[t]) -
                       ws_log_abs_jacobian(z_old[t], M.lo[t], M.hi[t], M.bound_kind[t]);
            }
        }
        lpr_out[i] = lpr;
        delta[i] = 0.0;
    }
}

// mode 0: out[i] += fold(segment)                      (score_logpdf)
// mode 1: out[i] += fold(segment | x_new) - fold(segment | current values)
__global__ void __launch_bounds__(WS_MOVE_BLOCK) ws_move_delta_kernel(const __grid_constant__ WsMoveParams M, int mode,
                                                                      const double* __restrict__ x_new, double* __restrict__ out) {
    extern __shared__ double ws_score_smem[];
    __shared__ WsDop dops[WS_FOLD_CHUNK];
    double* R = ws_score_smem + threadIdx.x;
    const WsScoreParams& S = M.score;
    const int64_t stride = (int64_t)gridDim.x * WS_MOVE_BLOCK;
    // CTA-uniform trip count (the fold has barriers); threads beyond n work on a clamped index and store nothing
    for (int64_t base = (int64_t)blockIdx.x * WS_MOVE_BLOCK; base < S.n; base += stride) {
        const bool live = base + threadIdx.x < S.n;
        const int64_t i = live ? base + threadIdx.x : S.n - 1;
        const uint64_t pid[1] = {(uint64_t)(S.particle_offset + i)};
        const int64_t idx[1] = {i};
        ws_score_load_planes<1>(S, R, idx);
        double s_old[1], s_new[1];
        ws_score_fold<1>(S, R, pid, s_old, dops);
        if (mode == 0) {
            if (live) out[i] += s_old[0] + S.konst;
        } else {
#pragma unroll
            for (int t = 0; t < WS_MOVE_MAX_D; ++t)
                if (t < M.d && M.target_reg[t] != 0xFF) R[(int)M.target_reg[t] * WS_MOVE_BLOCK] = x_new[(size_t)t * S.n + i];
            ws_score_fold<1>(S, R, pid, s_new, dops);
            if (live) out[i] += s_new[0] - s_old[0];
        }
    }
}

__global__ void __launch_bounds__(256) ws_move_accept_kernel(const __grid_constant__ WsMoveParams M, const double* __restrict__ x_new,
                                                             const double* __restrict__ lpr, const double* __restrict__ delta) {
    const int64_t n = M.score.n;
    unsigned long long accepted = 0ull;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const uint64_t particle = (uint64_t)(M.score.particle_offset + i);
        double u;
        if (M.rng.replay_u != nullptr) {
            u = M.rng.replay_u[M.replay_u_base + (int64_t)particle];
        } else {
            ws_u32x4 r = ws_philox4x32_10(particle, M.stream_uniform, M.rng.seed);
            u = ws_u01(r.x, r.y);
        }
        if (log(u) < lpr[i] + delta[i]) {
#pragma unroll
            for (int t = 0; t < WS_MOVE_MAX_D; ++t)
                if (t < M.d) M.target_ptr[t][i] = x_new[(size_t)t * n + i];
            ++accepted;
        }
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) accepted += __shfl_down_sync(0xffffffffu, accepted, dlt);
    if ((threadIdx.x & 31) == 0 && accepted != 0ull) atomicAdd(M.n_accept, accepted);
}

// ------------------------------------------------------------------------------------------
// autoRW moments.  pass 0: [sum w, sum w z_t];  pass 1: [-, -, sum w (z_i - m_i)(z_j - m_j) (j <= i)]
// Output layout per CTA: n_mom = 1 + d + d(d+1)/2 doubles.
// ------------------------------------------------------------------------------------------
#define WS_MOM_MAX (1 + WS_MOVE_MAX_D + WS_MOVE_MAX_D * (WS_MOVE_MAX_D + 1) / 2)

__global__ void __launch_bounds__(256) ws_move_moments_kernel(const __grid_constant__ WsMoveParams M, int64_t n, int pass,
                                                              double* __restrict__ partials) {
    const int d = M.d;
    const int n_mom = 1 + d + d * (d + 1) / 2;
    double acc[WS_MOM_MAX];
#pragma unroll
    for (int k = 0; k < WS_MOM_MAX; ++k) acc[k] = 0.0;
    double m = 0.0, S = 1.0;
    if (!M.w_uniform) {
        m = M.red->m;
        S = M.red->S;
    }
    const double rS = 1.0 / S;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const double w = M.w_uniform ? 1.0 : ws_div_pos(ws_exp_nonpos(M.logw[i] - m), S, rS);
        double z[WS_MOVE_MAX_D];
#pragma unroll
        for (int t = 0; t < WS_MOVE_MAX_D; ++t)
            if (t < d) z[t] = ws_to_unconstrained(M.target_ptr[t][i], M.lo[t], M.hi[t], M.bound_kind[t]);
        if (pass == 0) {
            acc[0] += w;
#pragma unroll
            for (int t = 0; t < WS_MOVE_MAX_D; ++t)
                if (t < d) acc[1 + t] += w * z[t];
        } else {
            int k = 1 + d;
#pragma unroll
            for (int a = 0; a < WS_MOVE_MAX_D; ++a) {
#pragma unroll
                for (int b = 0; b < WS_MOVE_MAX_D; ++b) {
                    if (a < d && b <= a) {
                        acc[k] += w * (z[a] - M.mean[a]) * (z[b] - M.mean[b]);
                        ++k;
                    }
                }
            }
        }
    }
    __shared__ double sc[8][WS_MOM_MAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < n_mom; ++k) {
        double v = acc[k];
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) v += __shfl_down_sync(0xffffffffu, v, dlt);
        if (lane == 0) sc[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < n_mom) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += sc[w][threadIdx.x];
        partials[(size_t)blockIdx.x * n_mom + threadIdx.x] = v;
    }
}

// ------------------------------------------------------------------------------------------
// exact distinct count: lock-free open-addressing set keyed by the value's bit pattern.
// Julia's unique() compares with isequal: every NaN is one value, -0.0 and 0.0 are two.
// ------------------------------------------------------------------------------------------
#define WS_SET_EMPTY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ unsigned long long ws_key_of(double v) {
    return (v != v) ? 0x7FF8000000000000ull : (unsigned long long)__double_as_longlong(v);
}
__device__ __forceinline__ unsigned long long ws_key_hash(unsigned long long key) {
    unsigned long long h = key * 0x9E3779B97F4A7C15ull;
    return h ^ (h >> 29);
}

// keys_are_bits: the input already holds canonical 64-bit keys (received from other ranks)
// part_*: when part_base != nullptr every NEW key is also appended to the list of its owner rank
// (owner = high hash bits mod n_parts), used by the sharded distinct count.
__global__ void __launch_bounds__(256) ws_unique_count_kernel(const double* __restrict__ plane, int64_t n,
                                                              unsigned long long* __restrict__ table, size_t mask,
                                                              unsigned long long* __restrict__ counter, int keys_are_bits,
                                                              unsigned long long* __restrict__ part_base, int64_t part_cap,
                                                              unsigned long long* __restrict__ part_count, int n_parts) {
    unsigned long long fresh = 0ull;
    const int64_t stride = (int64_t)gridDim.x * 256;
    const unsigned long long* bits = reinterpret_cast<const unsigned long long*>(plane);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = keys_are_bits ? bits[i] : ws_key_of(plane[i]);
        if (i > 0) {
            // resampled copies sit next to each other (ancestors are sorted): skip exact repeats
            const unsigned long long pkey = keys_are_bits ? bits[i - 1] : ws_key_of(plane[i - 1]);
            if (pkey == key) continue;
        }
        const unsigned long long h = ws_key_hash(key);
        size_t slot = (size_t)h & mask;
        while (true) {
            const unsigned long long old = atomicCAS(table + slot, WS_SET_EMPTY, key);
            if (old == WS_SET_EMPTY) {
                ++fresh;
                if (part_base != nullptr) {
                    const int owner = (int)((h >> 40) % (unsigned long long)n_parts);
                    const unsigned long long pos = atomicAdd(part_count + owner, 1ull);
                    part_base[(size_t)owner * part_cap + pos] = key;
                }
                break;
            }
            if (old == key) break;
            slot = (slot + 1) & mask;
        }
    }
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) fresh += __shfl_down_sync(0xffffffffu, fresh, dlt);
    if ((threadIdx.x & 31) == 0 && fresh != 0ull) atomicAdd(counter, fresh);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static int score_grid(int rows, int P, int64_t n, int sm_count) {
    const int smem = rows * P * WS_MOVE_BLOCK * (int)sizeof(double) + WS_FOLD_CHUNK * (int)sizeof(WsDop) + 1280;
    int per_sm = (227 * 1024) / smem;
    if (per_sm > 2048 / WS_MOVE_BLOCK) per_sm = 2048 / WS_MOVE_BLOCK;
    if (per_sm < 1) per_sm = 1;
    int64_t g = (n + (int64_t)P * WS_MOVE_BLOCK - 1) / ((int64_t)P * WS_MOVE_BLOCK);
    if (g > (int64_t)per_sm * sm_count) g = (int64_t)per_sm * sm_count;
    if (g < 1) g = 1;
    return (int)g;
}

// particles per thread of a fold: as many as keep at least four CTAs (16 warps) resident per SM
static int score_particles_per_thread(int rows) {
    const int row_bytes = WS_MOVE_BLOCK * (int)sizeof(double);
    if (rows * 4 * row_bytes <= 49 * 1024) return 4;  // + 6 KB of staged tape entries per CTA
    if (rows * 2 * row_bytes <= 49 * 1024) return 2;
    return 1;
}

cudaError_t ws_launch_score(const WsScoreParams& S, int sm_count, cudaStream_t s) {
    const int rows = S.n_regs;
    const int P = score_particles_per_thread(rows);
    const int smem = rows * P * WS_MOVE_BLOCK * (int)sizeof(double);
    const int g = score_grid(rows, P, S.n, sm_count);
    if (P == 4) ws_score_kernel<4><<<g, WS_MOVE_BLOCK, smem, s>>>(S);
    else if (P == 2) ws_score_kernel<2><<<g, WS_MOVE_BLOCK, smem, s>>>(S);
    else ws_score_kernel<1><<<g, WS_MOVE_BLOCK, smem, s>>>(S);
    return cudaGetLastError();
}

cudaError_t ws_launch_move(const WsMoveParams& M, int sm_count, cudaStream_t s) {
    const int rows = M.score.n_regs + M.d;  // score registers + proposed target values
    const int P = score_particles_per_thread(rows);
    const int smem = rows * P * WS_MOVE_BLOCK * (int)sizeof(double);
    const int g = score_grid(rows, P, M.score.n, sm_count);
    if (P == 4) ws_move_kernel<4><<<g, WS_MOVE_BLOCK, smem, s>>>(M);
    else if (P == 2) ws_move_kernel<2><<<g, WS_MOVE_BLOCK, smem, s>>>(M);
    else ws_move_kernel<1><<<g, WS_MOVE_BLOCK, smem, s>>>(M);
    return cudaGetLastError();
}

cudaError_t ws_launch_move_propose(const WsMoveParams& M, double* x_new, double* lpr, double* delta, int sm_count, cudaStream_t s) {
    int64_t g = (M.score.n + 255) / 256;
    if (g > (int64_t)sm_count * 8) g = (int64_t)sm_count * 8;
    if (g < 1) g = 1;
    ws_move_propose_kernel<<<(int)g, 256, 0, s>>>(M, x_new, lpr, delta);
    return cudaGetLastError();
}
cudaError_t ws_launch_move_delta(const WsMoveParams& M, int mode, const double* x_new, double* out, int sm_count, cudaStream_t s) {
    const int smem = M.score.n_regs * WS_MOVE_BLOCK * (int)sizeof(double);
    ws_move_delta_kernel<<<score_grid(M.score.n_regs, 1, M.score.n, sm_count), WS_MOVE_BLOCK, smem, s>>>(M, mode, x_new, out);
    return cudaGetLastError();
}
cudaError_t ws_launch_move_accept(const WsMoveParams& M, const double* x_new, const double* lpr, const double* delta, int sm_count,
                                  cudaStream_t s) {
    int64_t g = (M.score.n + 255) / 256;
    if (g > (int64_t)sm_count * 8) g = (int64_t)sm_count * 8;
    if (g < 1) g = 1;
    ws_move_accept_kernel<<<(int)g, 256, 0, s>>>(M, x_new, lpr, delta);
    return cudaGetLastError();
}

cudaError_t ws_launch_move_moments(const WsMoveParams& M, int64_t n, int pass, double* partials, int grid, cudaStream_t s) {
    ws_move_moments_kernel<<<grid, 256, 0, s>>>(M, n, pass, partials);
    return cudaGetLastError();
}

cudaError_t ws_launch_unique_count(const double* plane, int64_t n, unsigned long long* table, size_t slots,
                                   unsigned long long* counter, int sm_count, cudaStream_t s, int keys_are_bits,
                                   unsigned long long* part_base, int64_t part_cap, unsigned long long* part_count,
                                   int n_parts) {
    cudaError_t e = cudaMemsetAsync(table, 0xFF, sizeof(unsigned long long) * slots, s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    if (part_count != nullptr) {
        e = cudaMemsetAsync(part_count, 0, sizeof(unsigned long long) * n_parts, s);
        if (e != cudaSuccess) return e;
    }
    int64_t g = (n + 255) / 256;
    if (g > (int64_t)sm_count * 8) g = (int64_t)sm_count * 8;
    if (g < 1) g = 1;
    ws_unique_count_kernel<<<(int)g, 256, 0, s>>>(plane, n, table, slots - 1, counter, keys_are_bits, part_base, part_cap,
                                                   part_count, n_parts);
    return cudaGetLastError();
}

cudaError_t ws_move_kernels_init(int device) {
    (void)device;
    cudaError_t e = cudaFuncSetAttribute(ws_score_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ws_score_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ws_score_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ws_move_delta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ws_move_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ws_move_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ws_move_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}
