// ws_move.h — launch descriptors of the score-tape / Metropolis-Hastings kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ws_internal.h"

#define WS_MOVE_BLOCK 128
#define WS_SCORE_TEMPS 12        // registers [0, 12) of a score program are statement temporaries
#define WS_SCORE_MAX_REGS 200    // 200 * 128 * 8 B = 200 KB of shared memory at most
#define WS_SCORE_SEG_REGS 64     // register-file rows of one SEGMENT of a wide tape: 64 KB, three CTAs per SM (measured 200: 6.8 s, 96: 4.1 s, 64: 3.4 s, 28: 3.3 s on C4 J=512; env WSB200_SEG_REGS overrides)
#define WS_SCORE_MAX_LOADS (WS_SCORE_MAX_REGS - WS_SCORE_TEMPS)
#define WS_MOVE_MAX_D 8

// The score tape: the log-densities of every Sample / Observe / Weight executed so far, compiled
// to micro-ops over their own register space (device form of the score! fold,
// src/transformers.jl:39,77,139,193,243,297 + src/types.jl:198-206).
struct WsScoreParams {
    int64_t n;
    int64_t particle_offset;
    const WsOp* ops;  // device memory
    int32_t n_ops;    // prefix of the tape to fold (depth cut-off)
    int32_t n_regs;
    int32_t n_loads;
    int32_t pad;
    double konst;       // particle-independent part of the folded log-densities (added by ws_score_logpdf; cancels in a move)
    double* score_out;  // ws_score_logpdf only
    const double* load_ptr[WS_SCORE_MAX_LOADS];
    uint8_t load_reg[WS_SCORE_MAX_LOADS];  // already renumbered
    uint8_t reg_map[256];  // tape register -> row of this launch's register file (see compact_score_regs)
};

struct WsMoveParams {
    int32_t d;
    int32_t normals_target_major;  // replay order of the proposal normals (RW unbounded: target-major)
    int32_t w_uniform;             // autoRW moments: weights are all equal
    int32_t pad;
    double* target_ptr[WS_MOVE_MAX_D];
    uint8_t target_reg[WS_MOVE_MAX_D];  // register of the target inside the score program, 0xFF: none
    int32_t bound_kind[WS_MOVE_MAX_D];
    double lo[WS_MOVE_MAX_D], hi[WS_MOVE_MAX_D];
    double L[WS_MOVE_MAX_D * WS_MOVE_MAX_D];  // row-major d x d lower factor, stride d
    double mean[WS_MOVE_MAX_D];               // moments pass 1
    const double* logw;
    const WsReduceOut* red;
    WsRng rng;
    uint64_t stream_normals, stream_uniform;
    int64_t replay_n_base, replay_u_base, n_global;
    unsigned long long* n_accept;
    WsScoreParams score;
};

cudaError_t ws_move_kernels_init(int device);
cudaError_t ws_launch_score(const WsScoreParams& S, int sm_count, cudaStream_t s);
cudaError_t ws_launch_move(const WsMoveParams& M, int sm_count, cudaStream_t s);
cudaError_t ws_launch_move_propose(const WsMoveParams& M, double* x_new, double* lpr, double* delta, int sm_count, cudaStream_t s);
cudaError_t ws_launch_move_delta(const WsMoveParams& M, int mode, const double* x_new, double* out, int sm_count, cudaStream_t s);
cudaError_t ws_launch_move_accept(const WsMoveParams& M, const double* x_new, const double* lpr, const double* delta, int sm_count,
                                  cudaStream_t s);
cudaError_t ws_launch_move_moments(const WsMoveParams& M, int64_t n, int pass, double* partials, int grid, cudaStream_t s);
cudaError_t ws_launch_unique_count(const double* plane, int64_t n, unsigned long long* table, size_t slots,
                                   unsigned long long* counter, int sm_count, cudaStream_t s, int keys_are_bits = 0,
                                   unsigned long long* part_base = nullptr, int64_t part_cap = 0,
                                   unsigned long long* part_count = nullptr, int n_parts = 0);
