/* abi_smoke.c — a plain C consumer of include/wsb200.h (TEST INFRASTRUCTURE).
 *
 * Compiled with `gcc -std=c99 -Iinclude tests/host/abi_smoke.c -L.../lib -lwsb200` and nothing else: no Python, no
 * C++, no CUDA headers.  It proves that the header is self-sufficient for the binding a reference maintainer would
 * write: it runs the reference's benchmark model
 *     x ~ Normal(0, x0_std); for y in data:  x ~ Normal(a x, q);  y => Normal(x, r)          (lgssm1d.jl:18-24)
 * with Resample() after every `~` / `=>` (rewrites.jl:707-711) and a final `x << RW(step)` on replayed streams and
 * compares log-evidence, the particle column and the weights with what the oracle wrote into the fixture
 * (tests/golden/abi_smoke.bin, produced by tests/golden/make_golden.py from oracle/models.py).
 *
 * usage: abi_smoke <fixture.bin>       exit code 0 = parity within 1e-9 relative
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "wsb200.h"

#define CHECK(call)                                                                                      \
    do {                                                                                                 \
        int rc__ = (call);                                                                               \
        if (rc__ != WS_OK) {                                                                             \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, ws_last_error(ctx));                    \
            return 2;                                                                                    \
        }                                                                                                \
    } while (0)

static ws_tok tok_const(double v) {
    ws_tok t;
    memset(&t, 0, sizeof t);
    t.op = WS_TOK_CONST;
    t.val = v;
    return t;
}
static ws_tok tok_plane(int32_t col, int32_t comp) {
    ws_tok t;
    memset(&t, 0, sizeof t);
    t.op = WS_TOK_PLANE;
    t.col = col;
    t.comp = comp;
    return t;
}
static ws_tok tok_op(int32_t op) {
    ws_tok t;
    memset(&t, 0, sizeof t);
    t.op = op;
    return t;
}
static ws_expr expr_of(const ws_tok* toks, int32_t n) {
    ws_expr e;
    e.toks = toks;
    e.n = n;
    e.reserved = 0;
    return e;
}

static double* read_doubles(FILE* f, int64_t n) {
    double* p = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (p == NULL || fread(p, sizeof(double), (size_t)n, f) != (size_t)n) {
        fprintf(stderr, "fixture truncated\n");
        exit(3);
    }
    return p;
}

int main(int argc, char** argv) {
    ws_ctx* ctx = NULL;
    if (argc < 2) {
        fprintf(stderr, "usage: abi_smoke <fixture.bin>\n");
        return 3;
    }
    FILE* f = fopen(argv[1], "rb");
    if (f == NULL) {
        perror(argv[1]);
        return 3;
    }
    int64_t hdr[4]; /* n, T, n_normals, n_uniforms */
    if (fread(hdr, sizeof(int64_t), 4, f) != 4) return 3;
    const int64_t n = hdr[0], T = hdr[1];
    double* par = read_doubles(f, 6); /* a q r x0_std step ess_perc_min */
    double* obs = read_doubles(f, T);
    double* normals = read_doubles(f, hdr[2]);
    double* uniforms = read_doubles(f, hdr[3]);
    double* want_le = read_doubles(f, 1);
    double* want_x = read_doubles(f, n);
    double* want_w = read_doubles(f, n);
    int64_t want_counts[3]; /* resamples done, accepted proposals, depth */
    if (fread(want_counts, sizeof(int64_t), 3, f) != 3) return 3;
    fclose(f);
    const double a = par[0], q = par[1], r = par[2], x0 = par[3], step = par[4], ess_min = par[5];

    if (ws_abi_version() != WSB200_ABI_VERSION) {
        fprintf(stderr, "ABI version mismatch: library %d, header %d\n", ws_abi_version(), WSB200_ABI_VERSION);
        return 2;
    }
    if (ws_create(&ctx, n, 0, 1234, ess_min, WS_RESAMPLER_STRATIFIED) != WS_OK) {
        fprintf(stderr, "ws_create: %s\n", ws_last_error(NULL));
        return 2;
    }
    CHECK(ws_set_replay_normals(ctx, normals, hdr[2]));
    CHECK(ws_set_replay_uniforms(ctx, uniforms, hdr[3]));
    CHECK(ws_begin_run(ctx));

    int32_t xcol = -1;
    CHECK(ws_col_ensure(ctx, "x", 1, &xcol));
    ws_resample_info info;
    /* x ~ Normal(0.0, x0_std) ; Resample() (a no-op: nothing has been weighted) */
    {
        ws_tok mu[1], sg[1];
        mu[0] = tok_const(0.0);
        sg[0] = tok_const(x0);
        ws_expr emu = expr_of(mu, 1), esg = expr_of(sg, 1);
        CHECK(ws_sample_normal(ctx, xcol, 0, &emu, &esg));
        CHECK(ws_resample(ctx, &info));
        if (info.fired != 0) {
            fprintf(stderr, "Resample after an unweighted Sample must be a no-op\n");
            return 1;
        }
    }
    int64_t resamples = 0;
    for (int64_t t = 0; t < T; ++t) {
        /* x ~ Normal(a * x, q) : the mean is the postfix expression  a x *  */
        ws_tok mu[3], sg[1], ob[1], sr[1];
        mu[0] = tok_const(a);
        mu[1] = tok_plane(xcol, 0);
        mu[2] = tok_op(WS_TOK_MUL);
        sg[0] = tok_const(q);
        ws_expr emu = expr_of(mu, 3), esg = expr_of(sg, 1);
        CHECK(ws_sample_normal(ctx, xcol, 0, &emu, &esg));
        CHECK(ws_resample(ctx, &info));
        /* y => Normal(x, r) */
        ob[0] = tok_const(obs[t]);
        mu[0] = tok_plane(xcol, 0);
        sr[0] = tok_const(r);
        ws_expr eob = expr_of(ob, 1), emx = expr_of(mu, 1), esr = expr_of(sr, 1);
        CHECK(ws_observe_normal(ctx, &eob, &emx, &esr));
        CHECK(ws_resample(ctx, &info));
        if (info.fired != 1) {
            fprintf(stderr, "Resample after an Observe must fire\n");
            return 1;
        }
        resamples += info.resampled;
    }
    /* x << RW(step) */
    ws_move_spec spec;
    ws_move_info minfo;
    memset(&spec, 0, sizeof spec);
    int32_t tcol[1], tcomp[1];
    tcol[0] = xcol;
    tcomp[0] = 0;
    spec.n_targets = 1;
    spec.col = tcol;
    spec.comp = tcomp;
    spec.proposal = WS_PROPOSAL_RW;
    spec.has_bounds = 0;
    spec.step = step;
    spec.diversity = NAN;
    spec.target_depth = -1;
    CHECK(ws_move(ctx, &spec, &minfo));

    double le = 0.0, ess = 0.0;
    CHECK(ws_log_evidence(ctx, &le, &ess));
    double* x = (double*)malloc(sizeof(double) * (size_t)n);
    double* w = (double*)malloc(sizeof(double) * (size_t)n);
    CHECK(ws_col_download(ctx, xcol, x));
    CHECK(ws_weights_download(ctx, w));
    int64_t depth = 0;
    CHECK(ws_get_flags(ctx, NULL, NULL, &depth));

    int64_t bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (fabs(x[i] - want_x[i]) > 1e-9 * (1.0 + fabs(want_x[i]))) ++bad;
        if (fabs(w[i] - want_w[i]) > 1e-9 * (1.0 + fabs(want_w[i]))) ++bad;
    }
    const int le_ok = fabs(le - want_le[0]) <= 1e-9 * fabs(want_le[0]);
    printf("abi_smoke: n=%lld T=%lld log-evidence %.15g (oracle %.15g) mismatching values %lld, resamples %lld (oracle %lld), "
           "accepted %lld (oracle %lld), depth %lld (oracle %lld)\n",
           (long long)n, (long long)T, le, want_le[0], (long long)bad, (long long)resamples, (long long)want_counts[0],
           (long long)minfo.n_accepted, (long long)want_counts[1], (long long)depth, (long long)want_counts[2]);
    const int ok = le_ok && bad == 0 && resamples == want_counts[0] && minfo.n_accepted == want_counts[1] && depth == want_counts[2] &&
                   minfo.ran == 1;
    CHECK(ws_destroy(ctx));
    return ok ? 0 : 1;
}
