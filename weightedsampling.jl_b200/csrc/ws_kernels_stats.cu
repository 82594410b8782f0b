// ws_kernels_stats.cu — device side of describe(state) (src/utils.jl:183-289): weighted mean / std / min / max,
// 8-bin weighted histogram and the weighted median of StatsBase.quantile(v, Weights(w), 0.5) for one plane.
//
// The reference sorts all (value, weight) pairs on the host.  Here the median is found by a radix SELECT
// over the order-preserving 64-bit image of the doubles: eight passes of 8 bits, each a weighted histogram
// of the keys that share the prefix chosen so far.  Weights are the resampler's 2^61 fixed-point integers,
// so the histogram sums are exact and the selected element does not depend on the summation order.
//   pass W   w_i = exp(l_i - m)/S -> fixed point q_i (once per describe call, shared by all planes)
//   pass 1   sum w x, min, max, NaN flag, (smallest value with non-zero weight, its smallest weight)
//   pass 2   sum w (x - mean)^2, histogram over [min, max]
//   8 x (histogram of the next key byte | pick the byte where the running mass first exceeds h)
//   pass F   neighbours of the selected value: largest smaller value, smallest weight at the value
// All of it is HBM streaming work: 8 (x) + 8 (q) bytes per particle and pass.
#include <vector>
#include <string.h>
#include "ws_internal.h"
#include "ws_stats.h"

static int g_stats_sms = 148;

__device__ __forceinline__ unsigned long long ws_key_of(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ws_value_of_key(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

// ---- pass W: fixed-point weights ------------------------------------------------------------------
__global__ void __launch_bounds__(256) ws_stats_weights_kernel(const double* __restrict__ logw, const WsReduceOut* __restrict__ red,
                                                                int uniform, int64_t n, int64_t n_global,
                                                                unsigned long long* __restrict__ q) {
    double m = 0.0, S = 1.0;
    if (!uniform) {
        m = red->m;
        S = red->S;
    }
    const double rS = 1.0 / S;
    const double wu = 1.0 / (double)n_global;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const double w = uniform ? wu : ws_div_pos(ws_exp_nonpos(logw[i] - m), S, rS);
        unsigned long long v = 0ull;
        if (w > 0.0) v = (w >= 1.0) ? (1ull << 61) : __double2ull_rn(w * 2305843009213693952.0);
        q[i] = v;
    }
}

// ---- pass 1 -------------------------------------------------------------------------------------------
struct WsStat1 {
    double swx, sw, mn, mx, minv_nz, minw_nz;
    int has_nan, pad;
};

__device__ __forceinline__ void stat1_merge(WsStat1& a, const WsStat1& b) {
    a.swx += b.swx;
    a.sw += b.sw;
    a.mn = fmin(a.mn, b.mn);
    a.mx = fmax(a.mx, b.mx);
    a.has_nan |= b.has_nan;
    if (b.minv_nz < a.minv_nz || (b.minv_nz == a.minv_nz && b.minw_nz < a.minw_nz)) {
        a.minv_nz = b.minv_nz;
        a.minw_nz = b.minw_nz;
    }
}

__global__ void __launch_bounds__(256) ws_stats_pass1_kernel(const double* __restrict__ x, const unsigned long long* __restrict__ q,
                                                              int64_t n, WsStat1* __restrict__ partials) {
    __shared__ WsStat1 sm[8];
    WsStat1 a;
    a.swx = 0.0;
    a.sw = 0.0;
    a.mn = INFINITY;
    a.mx = -INFINITY;
    a.minv_nz = INFINITY;
    a.minw_nz = INFINITY;
    a.has_nan = 0;
    a.pad = 0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const double v = x[i];
        const double w = (double)q[i] * (1.0 / 2305843009213693952.0);
        if (v != v) a.has_nan = 1;
        a.swx += w * v;
        a.sw += w;
        a.mn = fmin(a.mn, v);
        a.mx = fmax(a.mx, v);
        if (w > 0.0 && (v < a.minv_nz || (v == a.minv_nz && w < a.minw_nz))) {
            a.minv_nz = v;
            a.minw_nz = w;
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        WsStat1 b;
        b.swx = __shfl_down_sync(0xffffffffu, a.swx, d);
        b.sw = __shfl_down_sync(0xffffffffu, a.sw, d);
        b.mn = __shfl_down_sync(0xffffffffu, a.mn, d);
        b.mx = __shfl_down_sync(0xffffffffu, a.mx, d);
        b.minv_nz = __shfl_down_sync(0xffffffffu, a.minv_nz, d);
        b.minw_nz = __shfl_down_sync(0xffffffffu, a.minw_nz, d);
        b.has_nan = __shfl_down_sync(0xffffffffu, a.has_nan, d);
        stat1_merge(a, b);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sm[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) stat1_merge(a, sm[w]);
        partials[blockIdx.x] = a;
    }
}

// ---- pass 2 -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ws_stats_pass2_kernel(const double* __restrict__ x, const unsigned long long* __restrict__ q,
                                                              int64_t n, double mean, double lo, double hi,
                                                              double* __restrict__ partials /* [grid][1 + 8] */) {
    __shared__ double sm[8][9];
    double acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0;
    const double scale = (hi > lo) ? 8.0 / (hi - lo) : 0.0;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const double v = x[i];
        const double w = (double)q[i] * (1.0 / 2305843009213693952.0);
        const double d = v - mean;
        acc[0] += w * d * d;
        int b = (int)((v - lo) * scale);  // clamp(searchsortedlast(range(lo, hi, 9), v), 1, 8) - 1
        b = b < 0 ? 0 : (b > 7 ? 7 : b);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k == b) acc[1 + k] += w;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        double v = acc[k];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane == 0) sm[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += sm[w][threadIdx.x];
        partials[(size_t)blockIdx.x * 9 + threadIdx.x] = v;
    }
}

// ---- radix select -------------------------------------------------------------------------------------
// state[0] = key prefix chosen so far (left-aligned), state[1] = fixed-point mass of all keys below the
// prefix range, state[2] = h (the fixed-point mass the running sum has to exceed), state[3] = mass of the
// selected byte's bin (after the last pass: total mass of the selected value), state[4] = 1 if the mass
// never exceeds h (the answer is the maximum)
__global__ void __launch_bounds__(256) ws_stats_hist_kernel(const double* __restrict__ x, const unsigned long long* __restrict__ q,
                                                             int64_t n, int pass, const unsigned long long* __restrict__ state,
                                                             unsigned long long* __restrict__ hist /* [256] */) {
    __shared__ unsigned long long h[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = lane; k < 256; k += 32) h[warp][k] = 0ull;
    __syncwarp();
    const unsigned long long prefix = state[0];
    const int shift = 56 - 8 * pass;
    const unsigned long long himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const unsigned long long key = ws_key_of(x[i]);
        const unsigned long long w = q[i];
        if ((key & himask) == prefix && w != 0ull) atomicAdd(&h[warp][(key >> shift) & 0xFFull], w);
    }
    __syncthreads();
    {
        unsigned long long s = 0ull;
        for (int w = 0; w < 8; ++w) s += h[w][threadIdx.x];
        if (s != 0ull) atomicAdd(hist + threadIdx.x, s);
    }
}

__global__ void ws_stats_pick_kernel(int pass, unsigned long long* __restrict__ state, unsigned long long* __restrict__ hist) {
    if (threadIdx.x != 0) return;
    unsigned long long below = state[1];
    const unsigned long long hh = state[2];
    int pick = -1;
    for (int b = 0; b < 256; ++b) {
        const unsigned long long c = hist[b];
        if (pick < 0 && c != 0ull && below + c > hh) {
            pick = b;
            state[3] = c;
        }
        if (pick < 0) below += c;
    }
    if (pick < 0) {
        state[4] = 1ull;  // the running mass never exceeds h: quantile() returns the maximum
        pick = 255;
    }
    state[0] |= (unsigned long long)pick << (56 - 8 * pass);
    state[1] = below;
    for (int b = 0; b < 256; ++b) hist[b] = 0ull;
}

// ---- pass F: neighbours of the selected value -----------------------------------------------------------
struct WsStatF {
    unsigned long long prev_key;  // largest key below the selected one among non-zero weights (0: none)
    unsigned long long minq_eq;   // smallest non-zero fixed-point weight at the selected value
};
__global__ void __launch_bounds__(256) ws_stats_final_kernel(const double* __restrict__ x, const unsigned long long* __restrict__ q,
                                                              int64_t n, const unsigned long long* __restrict__ state,
                                                              WsStatF* __restrict__ partials) {
    __shared__ WsStatF sm[8];
    const unsigned long long sel = state[0];
    unsigned long long pk = 0ull, mq = ~0ull;
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
        const unsigned long long w = q[i];
        if (w == 0ull) continue;
        const unsigned long long key = ws_key_of(x[i]);
        if (key < sel) pk = key > pk ? key : pk;
        else if (key == sel) mq = w < mq ? w : mq;
    }
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long a = __shfl_down_sync(0xffffffffu, pk, d), b = __shfl_down_sync(0xffffffffu, mq, d);
        pk = a > pk ? a : pk;
        mq = b < mq ? b : mq;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sm[warp].prev_key = pk;
        sm[warp].minq_eq = mq;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            pk = sm[w].prev_key > pk ? sm[w].prev_key : pk;
            mq = sm[w].minq_eq < mq ? sm[w].minq_eq : mq;
        }
        partials[blockIdx.x].prev_key = pk;
        partials[blockIdx.x].minq_eq = mq;
    }
}

// ---- host-side driver (called by ws_describe in ws_runtime.cu) -------------------------------------------
static int stats_grid(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = (int64_t)g_stats_sms * 8;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

cudaError_t ws_stats_init(int device) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess) g_stats_sms = prop.multiProcessorCount;
    return e;
}

cudaError_t ws_stats_weights(const double* logw, const WsReduceOut* red, int uniform, int64_t n, int64_t n_global,
                             unsigned long long* q, cudaStream_t s) {
    ws_stats_weights_kernel<<<stats_grid(n), 256, 0, s>>>(logw, red, uniform, n, n_global, q);
    return cudaGetLastError();
}

// scratch: device buffer of at least ws_stats_scratch_bytes(n); h_scratch: pinned host buffer of the same size
size_t ws_stats_scratch_bytes(int64_t n) {
    const size_t g = (size_t)stats_grid(n);
    return g * sizeof(WsStat1) + g * 9 * sizeof(double) + g * sizeof(WsStatF) + 256 * 8 + 8 * 8 + 256;
}

static void stat1_merge_host(WsStat1& a, const WsStat1& b) {
    a.swx += b.swx;
    a.sw += b.sw;
    a.mn = fmin(a.mn, b.mn);
    a.mx = fmax(a.mx, b.mx);
    a.has_nan |= b.has_nan;
    if (b.minv_nz < a.minv_nz || (b.minv_nz == a.minv_nz && b.minw_nz < a.minw_nz)) {
        a.minv_nz = b.minv_nz;
        a.minw_nz = b.minw_nz;
    }
}

cudaError_t ws_stats_plane(const double* x, const unsigned long long* q, int64_t n, void* d_scratch, void* h_scratch,
                           cudaStream_t s, WsPlaneStats* out, int* n_launches, const WsStatsComm* comm) {
    static_assert(sizeof(WsStat1) % 8 == 0, "WsStat1 travels as 64-bit words");
    const int g = stats_grid(n);
    char* dp = (char*)d_scratch;
    char* hp = (char*)h_scratch;
    WsStat1* d1 = (WsStat1*)dp;
    double* d2 = (double*)(dp + (size_t)g * sizeof(WsStat1));
    WsStatF* dF = (WsStatF*)((char*)d2 + (size_t)g * 9 * sizeof(double));
    unsigned long long* d_hist = (unsigned long long*)((char*)dF + (size_t)g * sizeof(WsStatF));
    unsigned long long* d_state = d_hist + 256;
    WsStat1* h1 = (WsStat1*)hp;
    double* h2 = (double*)(hp + (size_t)g * sizeof(WsStat1));
    WsStatF* hF = (WsStatF*)((char*)h2 + (size_t)g * 9 * sizeof(double));
    unsigned long long* h_state = (unsigned long long*)((char*)hF + (size_t)g * sizeof(WsStatF)) + 256;
    cudaError_t e;
    const double FX = 2305843009213693952.0;

    ws_stats_pass1_kernel<<<g, 256, 0, s>>>(x, q, n, d1);
    if ((e = cudaMemcpyAsync(h1, d1, sizeof(WsStat1) * g, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    WsStat1 a = h1[0];
    for (int k = 1; k < g; ++k) stat1_merge_host(a, h1[k]);
    if (comm != nullptr) {  // shard partials -> global, merged in rank order on every rank
        const size_t words = sizeof(WsStat1) / 8;
        std::vector<unsigned long long> all(words * (size_t)comm->nranks);
        if (comm->allgather_words(comm->ctx, reinterpret_cast<const unsigned long long*>(&a), words, all.data()) != 0) return cudaErrorUnknown;
        memcpy(&a, all.data(), sizeof(WsStat1));
        for (int r = 1; r < comm->nranks; ++r) {
            WsStat1 b;
            memcpy(&b, all.data() + words * (size_t)r, sizeof(WsStat1));
            stat1_merge_host(a, b);
        }
    }
    *n_launches = 1;
    out->min = a.mn;
    out->max = a.mx;
    out->mean = a.swx / a.sw;
    if (a.has_nan) {
        out->min = out->max = out->mean = out->std = out->median = NAN;
        for (int k = 0; k < 8; ++k) out->hist[k] = NAN;
        return cudaSuccess;
    }

    ws_stats_pass2_kernel<<<g, 256, 0, s>>>(x, q, n, out->mean, a.mn, a.mx, d2);
    if ((e = cudaMemcpyAsync(h2, d2, sizeof(double) * 9 * g, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;

    // weighted median: StatsBase.quantile(v, Weights(w), 0.5): h = p (wsum - w1) + w1, first k with S_k > h
    const double h = 0.5 * (a.sw - a.minw_nz) + a.minw_nz;
    unsigned long long st0[8] = {0ull, 0ull, (unsigned long long)floor(h * FX), 0ull, 0ull, 0ull, 0ull, 0ull};
    if ((e = cudaMemcpyAsync(d_state, st0, sizeof(st0), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(d_hist, 0, 256 * 8, s)) != cudaSuccess) return e;
    for (int pass = 0; pass < 8; ++pass) {
        ws_stats_hist_kernel<<<g, 256, 0, s>>>(x, q, n, pass, d_state, d_hist);
        if (comm != nullptr && comm->allreduce_u64_device(comm->ctx, d_hist, 256) != 0) return cudaErrorUnknown;  // exact integer sums
        ws_stats_pick_kernel<<<1, 32, 0, s>>>(pass, d_state, d_hist);
    }
    ws_stats_final_kernel<<<g, 256, 0, s>>>(x, q, n, d_state, dF);
    if ((e = cudaMemcpyAsync(hF, dF, sizeof(WsStatF) * g, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(h_state, d_state, 8 * 8, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    *n_launches += 1 + 16 + 1;

    double ss = 0.0;
    for (int k = 0; k < 8; ++k) out->hist[k] = 0.0;
    for (int b = 0; b < g; ++b) {
        ss += h2[(size_t)b * 9];
        for (int k = 0; k < 8; ++k) out->hist[k] += h2[(size_t)b * 9 + 1 + k];
    }
    if (comm != nullptr) {
        double v[9] = {ss};
        for (int k = 0; k < 8; ++k) v[1 + k] = out->hist[k];
        if (comm->allreduce_doubles(comm->ctx, v, 9) != 0) return cudaErrorUnknown;
        ss = v[0];
        for (int k = 0; k < 8; ++k) out->hist[k] = v[1 + k];
    }
    out->std = sqrt(ss / a.sw);

    unsigned long long pk = 0ull, mq = ~0ull;
    for (int b = 0; b < g; ++b) {
        pk = hF[b].prev_key > pk ? hF[b].prev_key : pk;
        mq = hF[b].minq_eq < mq ? hF[b].minq_eq : mq;
    }
    if (comm != nullptr) {
        const unsigned long long mine[2] = {pk, mq};
        std::vector<unsigned long long> all(2 * (size_t)comm->nranks);
        if (comm->allgather_words(comm->ctx, mine, 2, all.data()) != 0) return cudaErrorUnknown;
        for (int r = 0; r < comm->nranks; ++r) {
            pk = all[2 * r] > pk ? all[2 * r] : pk;
            mq = all[2 * r + 1] < mq ? all[2 * r + 1] : mq;
        }
    }
    const unsigned long long sel = h_state[0], below = h_state[1];
    auto val = [](unsigned long long k) {
        const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
        double v;
        memcpy(&v, &b, 8);
        return v;
    };
    const double vsel = val(sel);
    if (h_state[4] != 0ull) {
        out->median = a.mx;  // "out was initialized with maximum v"
    } else if (pk != 0ull && mq != ~0ull && below + mq > h_state[2]) {
        // the running mass crosses h at the FIRST element of the selected value's group: interpolate from the
        // previous element,  v_{k-1} + (h - S_{k-1}) / (S_k - S_{k-1}) (v_k - v_{k-1})
        const double Skold = (double)below / FX, wk = (double)mq / FX, vprev = val(pk);
        out->median = vprev + (h - Skold) / wk * (vsel - vprev);
    } else {
        out->median = vsel;
    }
    return cudaSuccess;
}


