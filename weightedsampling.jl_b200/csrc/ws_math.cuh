// ws_math.cuh — scalar building blocks shared by every wsb200 kernel.
//
// Everything here is a pure function of its arguments, marked WS_HD so that
//   * nvcc compiles it into the sm_100a kernels (the product path), and
//   * tests/host_math_check.cpp can compile the very same source with g++ and
//     compare it against the oracle without a GPU (test infrastructure only;
//     nothing in the shipped library runs these on the host).
//
// Reference formulas restated here (third-party, un-vendored):
//   Normal      rand = mu + sigma*z ; logpdf = -(zv^2 + log2pi)/2 - log(sigma)
//               (StatsFuns normlogpdf, used through src/default_kernels.jl:12-23,94)
//   Exponential rand = theta*e ; logpdf = log(1/theta) - x/theta, -Inf for x<0
//               (src/default_kernels.jl:87)
//   bound transforms / Jacobians: src/move_kernels.jl:37-85
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define WS_HD __host__ __device__ __forceinline__
#else
#define WS_HD inline
#endif

#define WS_LOG2PI 1.8378770664093453
#define WS_TWO_M53 1.1102230246251565e-16 /* 2^-53 */

// ---------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al., SC'11).  Counter layout used
// everywhere in wsb200:  (particle_lo, particle_hi, stream_lo, stream_hi),
// key = (seed_lo, seed_hi).  `particle` is the GLOBAL particle / slot index, so
// draws do not depend on grid shape, fusion window or the number of ranks.
// ---------------------------------------------------------------------------
struct ws_u32x4 {
    uint32_t x, y, z, w;
};

WS_HD void ws_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

WS_HD ws_u32x4 ws_philox4x32_10(uint64_t particle, uint64_t stream, uint64_t seed) {
    uint32_t c0 = (uint32_t)particle, c1 = (uint32_t)(particle >> 32);
    uint32_t c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        ws_mulhilo(0xD2511F53u, c0, hi0, lo0);
        ws_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    ws_u32x4 out;
    out.x = c0;
    out.y = c1;
    out.z = c2;
    out.w = c3;
    return out;
}

// 53-bit uniform in [0,1) from two 32-bit words (same support as Julia's rand()).
WS_HD double ws_u01(uint32_t hi, uint32_t lo) {
    uint64_t v = (((uint64_t)hi << 32) | (uint64_t)lo) >> 11;
    return (double)v * WS_TWO_M53;
}
// 53-bit uniform in (0,1] (safe argument for log).
WS_HD double ws_u01_open0(uint32_t hi, uint32_t lo) {
    uint64_t v = ((((uint64_t)hi << 32) | (uint64_t)lo) >> 11) + 1ull;
    return (double)v * WS_TWO_M53;
}

// One Philox block -> two independent standard normals (Box-Muller, FP64).
WS_HD void ws_randn2(uint64_t particle, uint64_t stream, uint64_t seed, double& z0, double& z1) {
    ws_u32x4 r = ws_philox4x32_10(particle, stream, seed);
    double u1 = ws_u01_open0(r.x, r.y);
    double u2 = ws_u01(r.z, r.w);
    double rad = sqrt(-2.0 * log(u1));
    double s, c;
#if defined(__CUDA_ARCH__)
    sincospi(2.0 * u2, &s, &c);
#else
    s = sin(6.283185307179586 * u2);
    c = cos(6.283185307179586 * u2);
#endif
    z0 = rad * c;
    z1 = rad * s;
}

// One Philox block -> two uniforms in [0,1).
WS_HD void ws_randu2(uint64_t particle, uint64_t stream, uint64_t seed, double& u0, double& u1) {
    ws_u32x4 r = ws_philox4x32_10(particle, stream, seed);
    u0 = ws_u01(r.x, r.y);
    u1 = ws_u01(r.z, r.w);
}

// Standard exponential variate e = -log(u), u in (0,1].
WS_HD double ws_randexp(uint64_t particle, uint64_t stream, uint64_t seed) {
    ws_u32x4 r = ws_philox4x32_10(particle, stream, seed);
    return -log(ws_u01_open0(r.x, r.y));
}

// ---------------------------------------------------------------------------
// log densities
// ---------------------------------------------------------------------------
WS_HD double ws_normal_logpdf(double x, double mu, double sigma) {
    double z;
    if (sigma == 0.0) {
        if (x == mu) {
            z = 0.0;  // zval(mu, 1, x)
        } else {
            z = (x - mu) / sigma;  // +-Inf
            sigma = 1.0;
        }
    } else {
        z = (x - mu) / sigma;
    }
    return -(z * z + WS_LOG2PI) / 2.0 - log(sigma);
}

WS_HD double ws_exponential_logpdf(double x, double theta) {
    double lam = 1.0 / theta;
    double z = log(lam) - lam * x;
    return x < 0.0 ? -INFINITY : z;
}

// ---------------------------------------------------------------------------
// MH bound transforms (src/move_kernels.jl:37-85).  kind: 0 none, 1 lower only,
// 2 upper only, 3 both.
// ---------------------------------------------------------------------------
WS_HD int ws_bound_kind(double lo, double hi) {
    bool flo = isfinite(lo), fhi = isfinite(hi);
    return (flo && fhi) ? 3 : (flo ? 1 : (fhi ? 2 : 0));
}
WS_HD double ws_to_unconstrained(double x, double lo, double hi, int kind) {
    switch (kind) {
        case 3: return log(x - lo) - log(hi - x);
        case 1: return log(x - lo);
        case 2: return log(hi - x);
        default: return x;
    }
}
WS_HD double ws_from_unconstrained(double z, double lo, double hi, int kind) {
    switch (kind) {
        case 3: return lo + (hi - lo) / (1.0 + exp(-z));
        case 1: return lo + exp(z);
        case 2: return hi - exp(z);
        default: return z;
    }
}
WS_HD double ws_log1pexp(double z) { return z > 0.0 ? z + log1p(exp(-z)) : log1p(exp(z)); }
WS_HD double ws_log_abs_jacobian(double z, double lo, double hi, int kind) {
    switch (kind) {
        case 3: return log(hi - lo) - ws_log1pexp(z) - ws_log1pexp(-z);
        case 1:
        case 2: return z;
        default: return 0.0;
    }
}

// ---------------------------------------------------------------------------
// Stratified / systematic slot uniforms in the reference's exact FP order
// (src/resampling.jl:39-41):  u_n = (n-1)*invN + r_n*invN,  n = 1..N.
// `n0` is the 0-based slot index (n-1).
// ---------------------------------------------------------------------------
// The explicit _rn intrinsics stop nvcc from contracting the expression into an
// FMA (Julia does not contract), so u_n is bit-identical to the reference's.
template <class I>
WS_HD double ws_slot_u(I n0, double r, double inv_n) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__dmul_rn((double)n0, inv_n), __dmul_rn(r, inv_n));
#else
    volatile double a = (double)n0 * inv_n;
    volatile double b = r * inv_n;
    return a + b;
#endif
}

// Fixed-point CDF: weights are accumulated as unsigned 64-bit integers with
// scale 2^63 so that prefix sums are exactly associative (deterministic for any
// scan order, tile shape or rank count).  See DESIGN.md "CDF arithmetic".
#define WS_FX_SCALE 9223372036854775808.0 /* 2^63 */
WS_HD uint64_t ws_weight_to_fx(double w) {
    // w in [0,1]; NaN / negative -> 0
    if (!(w > 0.0)) return 0ull;
    if (w >= 1.0) return 1ull << 63;
    double s = w * WS_FX_SCALE;
#if defined(__CUDA_ARCH__)
    return __double2ull_rn(s);
#else
    return (uint64_t)llrint(s);
#endif
}
WS_HD double ws_fx_to_double(uint64_t c) { return (double)c * (1.0 / WS_FX_SCALE); }

// Number of output slots whose uniform is <= C, i.e. F(C) = #{n0 in [0,N) : u(n0) <= C}
// for the (monotone) stratified / systematic grid u(n0) = n0*invN + r(n0)*invN.
// Offspring of particle m are exactly the slots  F(C_{m-1}) <= n0 < F(C_m), which
// restates icdf's  a_n = min{ m : C_m >= u_n }  (src/resampling.jl:13-26) per particle
// instead of per slot.  `r` is any callable slot -> uniform in [0,1).
template <class RFn>
WS_HD int64_t ws_count_slots_le(double C, int64_t N, double inv_n, RFn& r) {
    // Invariant wanted on exit: u(k-1) <= C (or k == 0) and u(k) > C (or k == N).  A uniform is only
    // drawn when the cheap bounds  fl(k*invN) <= u(k) <= fl(fl(k*invN) + invN)  cannot decide, which
    // leaves about one draw per call.  N < 2^31 (Int32 ancestors), so k fits an int.
    const double t = C * (double)N;
    int k = (t >= (double)N) ? (int)N : (t > 0.0 ? (int)t : 0);
    const int n = (int)N;
    bool advanced = false;
    while (k < n) {
        if (C < ws_slot_u(k, 0.0, inv_n)) break;  // u(k) >= fl(k*invN) > C
        if (ws_slot_u(k, r((int64_t)k), inv_n) <= C) {
            ++k;
            advanced = true;
        } else {
            break;
        }
    }
    if (!advanced) {
        while (k > 0) {
            if (C >= ws_slot_u(k - 1, 1.0, inv_n)) break;  // u(k-1) <= fl(fl((k-1)*invN) + invN) <= C
            if (ws_slot_u(k - 1, r((int64_t)(k - 1)), inv_n) > C) --k; else break;
        }
    }
    return (int64_t)k;
}
