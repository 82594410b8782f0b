/* ws_oracle.c — single-threaded C restatement of the reference's particle hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle/ref.py header): used by tests/ as a fast checker for large
 * N and by bench.py as the timed CPU baseline ("restated reference, not Julia": no Julia toolchain
 * exists in the image).  Parity with the Julia reference is UNPINNED at the bit
 * level (no golden vectors upstream); see oracle/ref.py for what pins it.
 *
 * Every loop below is one of the reference's N-length passes, in the same order and with the same
 * number of passes over memory as the Julia code performs (it allocates fresh arrays where Julia
 * does: w, us, indices), so that the timing is representative:
 *   exp_norm            src/resampling.jl:72-77     (max, exp, sum, divide: 4 passes, allocates w)
 *   ess_perc            src/resampling.jl:51-54
 *   logsumexp           src/resampling.jl:61-64     (2 passes)
 *   stratified_resample src/resampling.jl:35-43     (N rand(), allocates us)
 *   icdf                src/resampling.jl:13-26     (sequential two-pointer walk, allocates indices)
 *   resample!/_gather!  src/stores.jl:105-121       (one gather per column, ping-pong swap)
 *   Resample.apply!     src/transformers.jl:474-498
 *   Sample/Observe/Assign.apply!  src/transformers.jl:28-32,172-182,228-235
 * RNG: xoshiro256++ (Julia's default generator family) with a 128-layer ziggurat for normals
 * (Julia's randn is a ziggurat too); streams are NOT Julia's.
 * Build: see oracle/Makefile (-O2 -ffp-contract=off, no fast-math: Julia does not contract).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LOG2PI 1.8378770664093453

/* ---------------------------------------------------------------- RNG */
typedef struct { uint64_t s[4]; } rng_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t* r) {
    uint64_t* s = r->s;
    const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static void rng_seed(rng_t* r, uint64_t seed) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 4; ++i) {
        z += 0x9E3779B97F4A7C15ull;
        uint64_t x = z;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        r->s[i] = x ^ (x >> 31);
    }
}
static inline double rng_u01(rng_t* r) { return (double)(rng_next(r) >> 11) * 1.1102230246251565e-16; }

/* Marsaglia-Tsang ziggurat, 128 layers */
static double zig_w[128], zig_f[128];
static uint32_t zig_k[128];
static int zig_ready = 0;
static void zig_init(void) {
    const double m1 = 2147483648.0;
    double dn = 3.442619855899, tn = dn, vn = 9.91256303526217e-3;
    double q = vn / exp(-0.5 * dn * dn);
    zig_k[0] = (uint32_t)((dn / q) * m1);
    zig_k[1] = 0;
    zig_w[0] = q / m1;
    zig_w[127] = dn / m1;
    zig_f[0] = 1.0;
    zig_f[127] = exp(-0.5 * dn * dn);
    for (int i = 126; i >= 1; --i) {
        dn = sqrt(-2.0 * log(vn / dn + exp(-0.5 * dn * dn)));
        zig_k[i + 1] = (uint32_t)((dn / tn) * m1);
        tn = dn;
        zig_f[i] = exp(-0.5 * dn * dn);
        zig_w[i] = dn / m1;
    }
    zig_ready = 1;
}
static inline double rng_randn(rng_t* r) {
    for (;;) {
        const uint64_t bits = rng_next(r);
        const int32_t hz = (int32_t)(bits >> 32);
        const uint32_t iz = (uint32_t)bits & 127u;
        const uint32_t ahz = (uint32_t)(hz < 0 ? -(int64_t)hz : hz);
        if (ahz < zig_k[iz]) return hz * zig_w[iz];
        if (iz == 0) {
            double x, y;
            do {
                x = -log(1.0 - rng_u01(r)) * 0.2904764;
                y = -log(1.0 - rng_u01(r));
            } while (y + y < x * x);
            return hz > 0 ? 3.442619855899 + x : -3.442619855899 - x;
        }
        const double x = hz * zig_w[iz];
        if (zig_f[iz] + rng_u01(r) * (zig_f[iz - 1] - zig_f[iz]) < exp(-0.5 * x * x)) return x;
    }
}

/* ---------------------------------------------------------------- resampling.jl */
static double pairwise_sum(const double* x, int64_t n) { /* Julia's sum: pairwise, 1024 base case */
    if (n <= 1024) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += x[i];
        return s;
    }
    const int64_t h = n / 2;
    return pairwise_sum(x, h) + pairwise_sum(x + h, n - h);
}

void orc_exp_norm(const double* logw, int64_t n, double* w) {
    double m = logw[0];
    for (int64_t i = 1; i < n; ++i) if (logw[i] > m) m = logw[i];
    for (int64_t i = 0; i < n; ++i) w[i] = exp(logw[i] - m);
    const double s = pairwise_sum(w, n);
    for (int64_t i = 0; i < n; ++i) w[i] /= s;
}

double orc_ess_perc(const double* w, int64_t n) {
    /* sum(abs2, w): pairwise over the squares */
    double s;
    if (n <= 1024) {
        s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += w[i] * w[i];
    } else {
        double* t = (double*)malloc(sizeof(double) * (size_t)n);
        for (int64_t i = 0; i < n; ++i) t[i] = w[i] * w[i];
        s = pairwise_sum(t, n);
        free(t);
    }
    return 1.0 / ((double)n * s);
}

double orc_logsumexp(const double* logw, int64_t n) {
    double m = logw[0];
    for (int64_t i = 1; i < n; ++i) if (logw[i] > m) m = logw[i];
    double* t = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) t[i] = exp(logw[i] - m);
    const double s = pairwise_sum(t, n);
    free(t);
    return m + log(s);
}

void orc_stratified_us(const double* r, int64_t n, double* us) {
    const double inv_n = 1.0 / (double)n;
    for (int64_t i = 0; i < n; ++i) us[i] = (double)i * inv_n + r[i] * inv_n;
}

/* 0-based; returns the number of slots that ran past the end (reference: BoundsError), clamped to n-1 */
int64_t orc_icdf(const double* w, const double* us, int64_t n, int64_t* idx) {
    double s = w[0];
    int64_t m = 0, clamped = 0;
    for (int64_t k = 0; k < n; ++k) {
        while (s < us[k]) {
            if (m + 1 >= n) { ++clamped; break; }
            ++m;
            s += w[m];
        }
        idx[k] = m;
    }
    return clamped;
}

void orc_gather(const double* src, const int64_t* idx, int64_t n, double* dst) {
    for (int64_t i = 0; i < n; ++i) dst[i] = src[idx[i]];
}

double orc_normal_logpdf(double x, double mu, double sigma) {
    const double z = (x - mu) / sigma;
    return -(z * z + LOG2PI) / 2.0 - log(sigma);
}

/* ---------------------------------------------------------------- Resample.apply! on P columns */
typedef struct {
    int64_t n;
    int n_planes;
    double** front;
    double** back;
    double* weights;
    int resampled, weights_changed;
    double ess_perc_min;
    rng_t rng;
    int64_t n_resampled;
} orc_state;

static void orc_resample(orc_state* st) {
    if (!st->weights_changed) return;
    const int64_t n = st->n;
    double* w = (double*)malloc(sizeof(double) * (size_t)n);
    orc_exp_norm(st->weights, n, w);
    const double ess = orc_ess_perc(w, n);
    if (ess < st->ess_perc_min) {
        double* us = (double*)malloc(sizeof(double) * (size_t)n);
        const double inv_n = 1.0 / (double)n;
        for (int64_t i = 0; i < n; ++i) us[i] = (double)i * inv_n + rng_u01(&st->rng) * inv_n;
        int64_t* idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
        orc_icdf(w, us, n, idx);
        const double mean_logw = orc_logsumexp(st->weights, n) - log((double)n);
        for (int p = 0; p < st->n_planes; ++p) orc_gather(st->front[p], idx, n, st->back[p]);
        double** t = st->front; st->front = st->back; st->back = t;
        for (int64_t i = 0; i < n; ++i) st->weights[i] = mean_logw;
        st->resampled = 1;
        st->n_resampled++;
        free(us);
        free(idx);
    } else {
        st->resampled = 0;
    }
    st->weights_changed = 0;
    free(w);
}

/* ---------------------------------------------------------------- timed workloads
 * 2-D SSM bootstrap filter, filter-only form of examples/2D_ssm.jl:7-17:
 *   x .= x + v ; dv ~ MvNormal([0,0], 0.1 I) ; v .= v + dv ; o => MvNormal(x, 0.5 I) ; Resample()
 * Planes: x1 x2 v1 v2 dv1 dv2.  Each statement is its own pass (the reference materialises the
 * right-hand side of an Assign into a temporary first: src/rewrites.jl:167).
 * Returns log-evidence; out_mean[2] = weighted posterior mean of x.                               */
double orc_ssm2d_run(int64_t n, int64_t T, const double* obs /* T x 2 */, uint64_t seed, double ess_perc_min,
                     double* out_mean, int64_t* out_n_resampled) {
    if (!zig_ready) zig_init();
    orc_state st;
    memset(&st, 0, sizeof(st));
    st.n = n;
    st.n_planes = 6;
    st.ess_perc_min = ess_perc_min;
    rng_seed(&st.rng, seed);
    double* planes[12];
    for (int p = 0; p < 12; ++p) planes[p] = (double*)calloc((size_t)n, sizeof(double));
    double* front[6]; double* back[6];
    for (int p = 0; p < 6; ++p) { front[p] = planes[p]; back[p] = planes[6 + p]; }
    st.front = front; st.back = back;
    st.weights = (double*)calloc((size_t)n, sizeof(double));
    double* tmp = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) { st.front[2][i] = 1.0; }          /* v = [1, 0], x = [0, 0] */
    const double sd_dv = sqrt(0.1), sd_obs = sqrt(0.5);
    const double c0 = -(2.0 * LOG2PI + 2.0 * log(0.5)) / 2.0;          /* -(d log2pi + logdet)/2 */
    for (int64_t t = 0; t < T; ++t) {
        double **f = st.front;
        for (int c = 0; c < 2; ++c) {                                    /* x .= x + v */
            for (int64_t i = 0; i < n; ++i) tmp[i] = f[c][i] + f[2 + c][i];
            memcpy(f[c], tmp, sizeof(double) * (size_t)n);
        }
        for (int64_t i = 0; i < n; ++i) {                                /* dv ~ MvNormal(0, 0.1 I) */
            f[4][i] = sd_dv * rng_randn(&st.rng);
            f[5][i] = sd_dv * rng_randn(&st.rng);
        }
        for (int c = 0; c < 2; ++c) {                                    /* v .= v + dv */
            for (int64_t i = 0; i < n; ++i) tmp[i] = f[2 + c][i] + f[4 + c][i];
            memcpy(f[2 + c], tmp, sizeof(double) * (size_t)n);
        }
        const double o1 = obs[2 * t], o2 = obs[2 * t + 1];
        for (int64_t i = 0; i < n; ++i) {                                /* o => MvNormal(x, 0.5 I) */
            const double y1 = (o1 - f[0][i]) / sd_obs, y2 = (o2 - f[1][i]) / sd_obs;
            st.weights[i] += c0 - 0.5 * (y1 * y1 + y2 * y2);
        }
        st.weights_changed = 1;
        orc_resample(&st);
    }
    double* w = (double*)malloc(sizeof(double) * (size_t)n);
    orc_exp_norm(st.weights, n, w);
    double m1 = 0.0, m2 = 0.0;
    for (int64_t i = 0; i < n; ++i) { m1 += w[i] * st.front[0][i]; m2 += w[i] * st.front[1][i]; }
    if (out_mean) { out_mean[0] = m1; out_mean[1] = m2; }
    if (out_n_resampled) *out_n_resampled = st.n_resampled;
    const double le = orc_logsumexp(st.weights, n) - log((double)n);
    free(w); free(tmp); free(st.weights);
    for (int p = 0; p < 12; ++p) free(planes[p]);
    return le;
}

/* LGSSM-1D bootstrap filter, benchmarks/ssm/WeightedSampling/lgssm1d.jl:18-24 (the model the
 * reference's published numbers are for):  x ~ N(0, x0_std); per y: x ~ N(a x, q); y => N(x, r). */
double orc_lgssm1d_run(int64_t n, int64_t T, const double* ys, double a, double q, double r, double x0_std,
                       uint64_t seed, double ess_perc_min, double* out_mean, int64_t* out_n_resampled) {
    if (!zig_ready) zig_init();
    orc_state st;
    memset(&st, 0, sizeof(st));
    st.n = n;
    st.n_planes = 1;
    st.ess_perc_min = ess_perc_min;
    rng_seed(&st.rng, seed);
    double* front[1]; double* back[1];
    front[0] = (double*)calloc((size_t)n, sizeof(double));
    back[0] = (double*)calloc((size_t)n, sizeof(double));
    st.front = front; st.back = back;
    st.weights = (double*)calloc((size_t)n, sizeof(double));
    double* tmp = (double*)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) st.front[0][i] = 0.0 + x0_std * rng_randn(&st.rng);
    const double logr = log(r);
    (void)logr;
    for (int64_t t = 0; t < T; ++t) {
        double* x = st.front[0];
        for (int64_t i = 0; i < n; ++i) tmp[i] = a * x[i];               /* argument materialised first */
        for (int64_t i = 0; i < n; ++i) x[i] = tmp[i] + q * rng_randn(&st.rng);
        const double y = ys[t];
        for (int64_t i = 0; i < n; ++i) st.weights[i] += orc_normal_logpdf(y, x[i], r);
        st.weights_changed = 1;
        orc_resample(&st);
    }
    double* w = (double*)malloc(sizeof(double) * (size_t)n);
    orc_exp_norm(st.weights, n, w);
    double m1 = 0.0;
    for (int64_t i = 0; i < n; ++i) m1 += w[i] * st.front[0][i];
    if (out_mean) out_mean[0] = m1;
    if (out_n_resampled) *out_n_resampled = st.n_resampled;
    const double le = orc_logsumexp(st.weights, n) - log((double)n);
    free(w); free(tmp); free(st.weights); free(st.front[0]); free(st.back[0]);
    return le;
}

/* Resampling-only microbenchmark step (C5): exp_norm + ess + stratified + icdf + logsumexp + gather of
 * P planes + weight reset, on caller-supplied log-weights. Returns ESS%. */
double orc_resample_once(const double* logw, int64_t n, int n_planes, double** front, double** back, uint64_t seed,
                         int64_t* idx_out) {
    rng_t rng;
    rng_seed(&rng, seed);
    double* w = (double*)malloc(sizeof(double) * (size_t)n);
    orc_exp_norm(logw, n, w);
    const double ess = orc_ess_perc(w, n);
    double* us = (double*)malloc(sizeof(double) * (size_t)n);
    const double inv_n = 1.0 / (double)n;
    for (int64_t i = 0; i < n; ++i) us[i] = (double)i * inv_n + rng_u01(&rng) * inv_n;
    orc_icdf(w, us, n, idx_out);
    (void)orc_logsumexp(logw, n);
    for (int p = 0; p < n_planes; ++p) orc_gather(front[p], idx_out, n, back[p]);
    free(w);
    free(us);
    return ess;
}
