// tests/host/harness.cpp — CPU test harness for the statement lowering (TEST INFRASTRUCTURE ONLY).
//
// It runs the SAME lowering (csrc/ws_lowering.h) and the SAME micro-op interpreter
// (csrc/ws_vm.cuh, host instantiation) that the CUDA runtime uses, over host arrays, so that
// `pytest -m "not gpu"` can check register allocation, op fusion, replay indexing and the score tape
// against the oracle without a GPU.  It is never part of the shipped library: libwsb200.so has no
// CPU path.
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>
#include "../../weightedsampling.jl_b200/csrc/ws_lowering.h"
#include "../../weightedsampling.jl_b200/csrc/ws_vm_sl.cuh"
#include "../../weightedsampling.jl_b200/csrc/ws_exchange.h"

using wsl::Plane;
using wsl::Program;

struct HH {
    int64_t n;
    std::vector<std::vector<std::vector<double>>> cols;  // [col][comp][i]
    std::vector<double> logw;
    Program win, score;
    uint64_t next_stream = 1, seed = 0;
    int64_t cur_n = 0, cur_u = 0, cur_e = 0, cur_v = 0;
    std::vector<double> rn, ru, re, rv;
    std::vector<int32_t> tape_end;
    std::vector<double> tape_const;  // Program::acc_const after each tape entry
    int n_flush = 0;
    std::string err;
};

static void reset_win(HH* h) {
    h->win = Program();
    h->win.max_regs = 64;
    h->win.max_ops = 96;
    h->win.max_io = 24;
}

extern "C" {

HH* hh_new(int64_t n, uint64_t seed) {
    HH* h = new HH();
    h->n = n;
    h->seed = seed;
    h->logw.assign((size_t)n, 0.0);
    reset_win(h);
    h->score = Program();
    h->score.score_mode = true;
    h->score.n_temp_slots = 12;
    h->score.max_regs = 200;
    h->score.max_ops = 1 << 30;
    h->score.max_io = 1 << 30;
    return h;
}
void hh_free(HH* h) { delete h; }
const char* hh_error(HH* h) { return h->err.c_str(); }
int hh_col(HH* h, int width) {
    h->cols.push_back(std::vector<std::vector<double>>(width, std::vector<double>((size_t)h->n, 0.0)));
    return (int)h->cols.size() - 1;
}
void hh_set_plane(HH* h, int col, int comp, const double* v) { memcpy(h->cols[col][comp].data(), v, sizeof(double) * h->n); }
void hh_set_replay(HH* h, const double* n, int64_t nn, const double* u, int64_t nu, const double* e, int64_t ne) {
    h->rn.assign(n, n + nn);
    h->ru.assign(u, u + nu);
    h->re.assign(e, e + ne);
}

static void run_program(HH* h, Program& p, std::vector<double>* acc_out, int n_ops) {
    WsRng rng;
    rng.seed = h->seed;
    rng.replay_n = h->rn.empty() ? nullptr : h->rn.data();
    rng.replay_u = h->ru.empty() ? nullptr : h->ru.data();
    rng.replay_e = h->re.empty() ? nullptr : h->re.data();
    rng.replay_v = h->rv.empty() ? nullptr : h->rv.data();
    std::vector<double> R((size_t)std::max(1, p.high_water));
    for (int64_t i = 0; i < h->n; ++i) {
        for (auto& ld : p.loads) R[ld.second] = h->cols[ld.first.col][ld.first.comp][(size_t)i];
        double acc[1] = {0.0};
        const uint64_t pid[1] = {(uint64_t)i};
        for (int pc = 0; pc < n_ops; ++pc) ws_vm_exec<1, 1>(p.ops[pc], R.data(), acc, rng, pid);
        for (auto& d : p.dirty) h->cols[d.col][d.comp][(size_t)i] = R[p.plane_reg[d]];
        if (acc_out) (*acc_out)[(size_t)i] = acc[0];
    }
}

int hh_flush(HH* h) {
    if (h->win.ops.empty() && h->win.dirty.empty()) return 0;
    if (h->win.overflow) {
        h->err = "window overflow";
        return -5;
    }
    std::vector<double> acc((size_t)h->n, 0.0);
    run_program(h, h->win, &acc, (int)h->win.ops.size());
    if (h->win.has_acc)
        for (int64_t i = 0; i < h->n; ++i) h->logw[(size_t)i] += acc[(size_t)i];
    h->n_flush++;
    reset_win(h);
    return 0;
}
int hh_n_flush(HH* h) { return h->n_flush; }
// debugging aid: the queued window as text (op dst a b c imm k0 k1 k2 per line)
int hh_dump_window(HH* h, char* buf, int cap) {
    std::string out;
    char line[256];
    for (auto& ld : h->win.loads) {
        snprintf(line, sizeof line, "load  r%d <- col %d comp %d\n", ld.second, ld.first.col, ld.first.comp);
        out += line;
    }
    for (auto& o : h->win.ops) {
        snprintf(line, sizeof line, "op %2u dst %3u a %3u b %3u c %3u imm %u k0 %.17g k1 %.17g k2 %.17g\n", o.w0 & 0xFFu,
                 (o.w0 >> 8) & 0xFFu, (o.w0 >> 16) & 0xFFu, (o.w0 >> 24) & 0xFFu, o.w1 & 0xFFu, o.w1 >> 8, o.k0, o.k1, o.k2);
        out += line;
    }
    for (auto& d : h->win.dirty) {
        snprintf(line, sizeof line, "store col %d comp %d <- r%d\n", d.col, d.comp, h->win.plane_reg[d]);
        out += line;
    }
    snprintf(buf, cap, "%s", out.c_str());
    return (int)out.size();
}
// which straight-line signature (csrc/ws_vm_sl.cuh) the queued window matches: its index, or -1 (interpreter).
// The view is filled exactly as flush_window (ws_runtime.cu) fills WsVmProgram.
int hh_window_signature(HH* h) {
    struct View {
        int n_ops, n_loads, n_stores, n_expect;
        uint8_t load_reg[64], store_reg[64];
        std::vector<WsOp> ops;
    } v;
    const Program& w = h->win;
    if (w.loads.size() > 64 || w.dirty.size() > 64) return -1;
    v.n_ops = (int)w.ops.size();
    v.n_loads = (int)w.loads.size();
    v.n_stores = (int)w.dirty.size();
    v.n_expect = 0;
    for (int k = 0; k < v.n_loads; ++k) v.load_reg[k] = (uint8_t)w.loads[k].second;
    for (int k = 0; k < v.n_stores; ++k) v.store_reg[k] = (uint8_t)w.plane_reg.at(w.dirty[k]);
    v.ops = w.ops;
    return ws_sl_find(v);
}
int hh_window_ops(HH* h) { return (int)h->win.ops.size(); }
int hh_window_regs(HH* h) { return h->win.high_water; }
int hh_window_loads(HH* h) { return (int)h->win.loads.size(); }
int hh_window_stores(HH* h) { return (int)h->win.dirty.size(); }

static wsl::RngCursor cursor(HH* h) { return wsl::RngCursor{&h->next_stream, &h->cur_n, &h->cur_u, &h->cur_e, h->n, &h->cur_v}; }
static int finish(HH* h, Program& p) {
    if (!p.error.empty()) {
        h->err = p.error;
        return -1;
    }
    p.end_statement();
    return 0;
}
static void tape(HH* h) {
    h->score.end_statement();
    h->tape_end.push_back((int32_t)h->score.ops.size());
    h->tape_const.push_back(h->score.acc_const);
}

int hh_assign(HH* h, int col, int comp, const ws_expr* rhs) {
    wsl::stmt_assign(h->win, Plane{col, comp}, *rhs);
    return finish(h, h->win);
}
int hh_assign_vec(HH* h, int col, int d, const ws_expr* rhs) {
    wsl::stmt_assign_vec(h->win, col, d, rhs);
    return finish(h, h->win);
}
int hh_sample_normal(HH* h, int col, int comp, const ws_expr* mu, const ws_expr* sigma) {
    wsl::RngCursor rc = cursor(h);
    wsl::stmt_sample_normal(h->win, rc, Plane{col, comp}, *mu, *sigma);
    wsl::score_sample_normal(h->score, Plane{col, comp}, *mu, *sigma);
    tape(h);
    return finish(h, h->win);
}
int hh_sample_exponential(HH* h, int col, int comp, const ws_expr* theta) {
    wsl::RngCursor rc = cursor(h);
    wsl::stmt_sample_exponential(h->win, rc, Plane{col, comp}, *theta);
    wsl::score_sample_exponential(h->score, Plane{col, comp}, *theta);
    tape(h);
    return finish(h, h->win);
}
int hh_sample_mvnormal(HH* h, int col, int d, const ws_expr* mu, const double* cov) {
    std::vector<double> L, Linv;
    double c0;
    if (!wsl::mvnormal_factors(d, cov, L, Linv, c0)) return -7;
    wsl::RngCursor rc = cursor(h);
    wsl::stmt_sample_mvnormal(h->win, rc, col, d, mu, L);
    wsl::score_sample_mvnormal(h->score, col, d, mu, Linv, c0);
    tape(h);
    return finish(h, h->win);
}
int hh_observe_normal(HH* h, const ws_expr* obs, const ws_expr* mu, const ws_expr* sigma) {
    wsl::stmt_observe_normal(h->win, *obs, *mu, *sigma);
    wsl::stmt_observe_normal(h->score, *obs, *mu, *sigma);
    tape(h);
    return finish(h, h->win);
}
int hh_observe_exponential(HH* h, const ws_expr* obs, const ws_expr* theta) {
    wsl::stmt_observe_exponential(h->win, *obs, *theta);
    wsl::stmt_observe_exponential(h->score, *obs, *theta);
    tape(h);
    return finish(h, h->win);
}
int hh_observe_mvnormal(HH* h, int d, const ws_expr* obs, const ws_expr* mu, const double* cov) {
    std::vector<double> L, Linv;
    double c0;
    if (!wsl::mvnormal_factors(d, cov, L, Linv, c0)) return -7;
    wsl::stmt_observe_mvnormal(h->win, d, obs, mu, Linv, c0);
    wsl::stmt_observe_mvnormal(h->score, d, obs, mu, Linv, c0);
    tape(h);
    return finish(h, h->win);
}
int hh_weight_expr(HH* h, const ws_expr* term) {
    wsl::stmt_weight_expr(h->win, *term);
    wsl::stmt_weight_expr(h->score, *term);
    tape(h);
    return finish(h, h->win);
}
int hh_sample_expr(HH* h, int col, int comp, const ws_expr* sampler, const ws_expr* weighter, const ws_expr* logpdf) {
    wsl::RngCursor rc = cursor(h);
    wsl::stmt_sample_expr(h->win, rc, Plane{col, comp}, *sampler, weighter);
    if (logpdf) {
        wsl::stmt_weight_expr(h->score, *logpdf);
        tape(h);
    }
    return finish(h, h->win);
}
int hh_importance_normal(HH* h, int col, int comp, double pm, double ps, double tm, double ts) {
    wsl::RngCursor rc = cursor(h);
    wsl::stmt_importance_normal(h->win, rc, Plane{col, comp}, pm, ps, tm, ts);
    return finish(h, h->win);
}
void hh_get_plane(HH* h, int col, int comp, double* out) {
    hh_flush(h);
    memcpy(out, h->cols[col][comp].data(), sizeof(double) * h->n);
}
void hh_get_logw(HH* h, double* out) {
    hh_flush(h);
    memcpy(out, h->logw.data(), sizeof(double) * h->n);
}
int hh_tape_len(HH* h) { return (int)h->tape_end.size(); }
int hh_score_ops(HH* h) { return (int)h->score.ops.size(); }
// fold the first n_entries tape entries at the current column values
int hh_score(HH* h, int n_entries, double* out) {
    hh_flush(h);
    if (h->score.overflow) return -5;
    if (n_entries > (int)h->tape_end.size()) return -1;
    const int n_ops = n_entries <= 0 ? 0 : h->tape_end[(size_t)n_entries - 1];
    std::vector<double> acc((size_t)h->n, 0.0);
    Program& p = h->score;
    std::vector<Plane> saved_dirty = p.dirty;
    run_program(h, p, &acc, n_ops);
    const double konst = n_entries <= 0 ? 0.0 : h->tape_const[(size_t)n_entries - 1];
    for (auto& v : acc) v += konst;
    memcpy(out, acc.data(), sizeof(double) * h->n);
    return 0;
}
// The same fold the way the move / score kernels run it (ws_kernels_move.cu: ws_score_fold): registers renumbered to
// the rows the prefix touches, entries unpacked to WsDop, chunks of 128 entries, runs of squared-residual entries over the
// same registers executed by ws_vm_exec_sqlin2_run.  Must equal hh_score bit for bit.  *n_runs / *n_rows report what
// the compaction and the run detection found.
int hh_score_device_order(HH* h, int n_entries, double* out, int* n_runs, int* n_rows) {
    hh_flush(h);
    if (h->score.overflow) return -5;
    if (n_entries > (int)h->tape_end.size()) return -1;
    const int n_ops = n_entries <= 0 ? 0 : h->tape_end[(size_t)n_entries - 1];
    Program& p = h->score;
    std::vector<uint8_t> keep;
    for (auto& ld : p.loads) keep.push_back((uint8_t)ld.second);
    uint8_t map[256];
    const int rows = std::max(1, ws_compact_regs(p.ops.data(), (size_t)n_ops, keep.data(), (int)keep.size(), map));
    const int CHUNK = 128;
    std::vector<WsDop> dops((size_t)n_ops);
    int runs = 0;
    for (int i = 0; i < n_ops; ++i) {
        dops[(size_t)i] = ws_decode_op<1, 1>(p.ops[(size_t)i], map);
        dops[(size_t)i].op |= 1u << 8;
    }
    for (int i = 0; i < n_ops; ++i) {
        const bool cont = (i % CHUNK) != 0 && ws_run_continues(p.ops[(size_t)i - 1], p.ops[(size_t)i]);
        if (cont) continue;
        int len = 1;
        while (i + len < n_ops && ((i + len) % CHUNK) != 0 && ws_run_continues(p.ops[(size_t)(i + len) - 1], p.ops[(size_t)(i + len)])) ++len;
        if (len > 1) {
            dops[(size_t)i].op = (dops[(size_t)i].op & 0xFFu) | ((uint32_t)len << 8);
            ++runs;
        }
    }
    WsRng none;
    none.seed = 0;
    none.replay_n = none.replay_u = none.replay_e = none.replay_v = nullptr;
    std::vector<double> R((size_t)rows);
    const double konst = n_entries <= 0 ? 0.0 : h->tape_const[(size_t)n_entries - 1];
    for (int64_t i = 0; i < h->n; ++i) {
        for (auto& ld : p.loads) R[map[ld.second]] = h->cols[ld.first.col][ld.first.comp][(size_t)i];
        double acc[1] = {0.0};
        const uint64_t pid[1] = {(uint64_t)i};
        for (int k = 0; k < n_ops;) {
            const int len = (int)(dops[(size_t)k].op >> 8);
            if (len > 1) ws_vm_exec_sqlin2_run<1, 1>(dops.data() + k, len, R.data(), acc);
            else ws_vm_exec_d<1, 1>(dops[(size_t)k], R.data(), acc, none, pid);
            k += len;
        }
        out[(size_t)i] = acc[0] + konst;
    }
    if (n_runs) *n_runs = runs;
    if (n_rows) *n_rows = rows;
    return 0;
}
// The exchange plan of a sharded resampling step (csrc/ws_exchange.h), rank r's view.  out: per rank d
// [send_off, send_cnt, recv_off, recv_cnt, spare_pos, push_offset_lazy, push_offset_eager] (7 x R), then
// [fs, fe, remote_send, remote_recv, total_remote, fits].
void hh_exchange_plan(const int32_t* bnd, int R, int r, int64_t n_global, int64_t* out) {
    const WsExchangePlan p = ws_exchange_plan(bnd, R, r, n_global);
    for (int d = 0; d < R; ++d) {
        int64_t* o = out + 7 * d;
        o[0] = p.send_off[(size_t)d];
        o[1] = p.send_cnt[(size_t)d];
        o[2] = p.recv_off[(size_t)d];
        o[3] = p.recv_cnt[(size_t)d];
        o[4] = p.spare_pos[(size_t)d];
        o[5] = d == r ? -1 : ws_push_offset(bnd, R, r, d, n_global, true);
        o[6] = d == r ? -1 : ws_push_offset(bnd, R, r, d, n_global, false);
    }
    int64_t* t = out + 7 * R;
    t[0] = p.fs;
    t[1] = p.fe;
    t[2] = p.remote_send;
    t[3] = p.remote_recv;
    t[4] = p.total_remote;
    t[5] = p.fits ? 1 : 0;
}
// Philox / Box-Muller / slot-count building blocks of ws_math.cuh
void hh_set_replay_variates(HH* h, const double* v, int64_t nv) { h->rv.assign(v, v + nv); }
// the rejection samplers of ws_math.cuh: n draws with particle = first + i
void hh_rand_gamma(double a, uint64_t first, uint64_t stream, uint64_t seed, double* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = ws_rand_gamma(a, first + (uint64_t)i, stream, seed);
}
void hh_rand_poisson(double lam, uint64_t first, uint64_t stream, uint64_t seed, double* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = ws_rand_poisson(lam, first + (uint64_t)i, stream, seed);
}
// fixed-point exponential spacings of slots 0 .. n_all-1 (multinomial resampling without a sort)
void hh_spacings(int64_t n_all, int mn_shift, uint64_t seed, uint64_t stream, uint64_t* out) {
    for (int64_t k = 0; k < n_all; ++k) out[k] = ws_spacing_of_slot((uint64_t)k, mn_shift, seed, stream);
}
void hh_randn2(uint64_t particle, uint64_t stream, uint64_t seed, double* out2) { ws_randn2(particle, stream, seed, out2[0], out2[1]); }
void hh_philox(uint64_t particle, uint64_t stream, uint64_t seed, uint32_t* out4) {
    ws_u32x4 r = ws_philox4x32_10(particle, stream, seed);
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
// the custom FP64 routines of ws_math.cuh (host instantiation)
void hh_exp_nonpos(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ws_exp_nonpos(x[i]); }
void hh_log_pos(const double* x, const int32_t* kb, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ws_log_pos(x[i], kb[i]); }
void hh_sqrt_pos(const double* x, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ws_sqrt_pos(x[i]); }
void hh_div_pos(const double* a, const double* b, double* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = ws_div_pos(a[i], b[i], 1.0 / b[i]); }
void hh_sincos_octant(const double* f, double* s, double* c, int64_t n) { for (int64_t i = 0; i < n; ++i) ws_sincos_octant(f[i], s[i], c[i]); }
void hh_box_muller(const uint64_t* w1, const uint64_t* w2, double* z0, double* z1, int64_t n) { for (int64_t i = 0; i < n; ++i) ws_box_muller(w1[i], w2[i], z0[i], z1[i]); }
// F(C) for every C in cs against the stratified grid built from r (replay)
struct RArr {
    const double* r;
    double operator()(int64_t k) { return r[k]; }
};
void hh_count_slots_le(const double* cs, int64_t n_c, const double* r, int64_t n, int64_t* out) {
    RArr ra{r};
    const double inv_n = 1.0 / (double)n;
    for (int64_t i = 0; i < n_c; ++i) out[i] = ws_count_slots_le(cs[i], n, inv_n, ra);
}
// The spare-row ring of the sharded genealogy (csrc/ws_exchange.h): replay a sequence of operations on one ring.
// ops[k] = (kind, a, b): kind 0 = allocate a rows for event b -> out[k] = start or -1 (a failed allocation changes nothing);
// kind 1 = release the regions of events <= a -> out[k] = regions left.
void hh_spare_ring(int64_t cap, const int64_t* ops, int n_ops, int64_t* out) {
    WsSpareRing ring;
    for (int k = 0; k < n_ops; ++k) {
        const int64_t kind = ops[3 * k], a = ops[3 * k + 1], b = ops[3 * k + 2];
        if (kind == 0) {
            const int64_t s = ws_spare_ring_peek(ring, cap, a);
            if (s >= 0) ring.push_back(WsSpareRegion{b, s, a});
            out[k] = s;
        } else {
            ws_spare_ring_release(ring, a);
            out[k] = (int64_t)ring.size();
        }
    }
}
// The mailbox protocol of csrc/ws_mailbox.cuh, restated for R host threads over std::atomic words (the device code uses
// st/ld.relaxed.sys on peer-mapped memory; the layout, the word format and the double-buffering are the same):
//   box[q] = rank q's mailbox, [2 parities][R sources][cap] words; a message word = (seq << 32) | 32 payload bits;
//   exchange `seq`: every rank stores its words into every box at [(seq & 1) * R + me], then reads its own box until each
//   word shows `seq`.
// Every rank runs K exchanges back to back with random pauses between and inside them (skewed ranks); a rank checks every
// payload it receives against what the sender must have sent for THAT exchange.  Returns the number of wrong payloads,
// or -1 if a rank waited longer than `spin_limit` polls (a word was overwritten before it was read: the reader would
// wait for ever) — both must be 0 / never for the argument "nobody can be two exchanges ahead of anybody" to hold.
// `buffers` = 2 is the protocol; 1 (no double-buffering) is the negative control the test also runs, with `slow_rank`
// reading slowly: the others finish the exchange (they only need its message, which it sent first), start the next one
// and overwrite words it has not read yet.
}  // extern "C"
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>
extern "C" {
static inline uint32_t hh_mbox_payload(int src, uint32_t seq, int w) { return (uint32_t)(src * 2654435761u) ^ (seq * 40503u) ^ ((uint32_t)w * 97u + 13u); }
int64_t hh_mailbox_protocol(int R, int K, int n_words, int buffers, uint64_t seed, int64_t spin_limit, int slow_rank) {
    const int ll = 2 * n_words, cap = ll;
    std::vector<std::unique_ptr<std::atomic<unsigned long long>[]>> box((size_t)R);
    for (auto& b : box) {
        b.reset(new std::atomic<unsigned long long>[(size_t)buffers * R * cap]);
        for (size_t i = 0; i < (size_t)buffers * R * cap; ++i) b[i].store(0ull, std::memory_order_relaxed);
    }
    std::atomic<int64_t> wrong{0};
    std::atomic<int> dead{0};
    auto rank_fn = [&](int me) {
        uint64_t s = seed * 6364136223846793005ull + (uint64_t)me * 1442695040888963407ull + 1ull;
        auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
        auto pause = [&]() {
            const int k = (int)(rnd() % 64);
            if (k < 8) std::this_thread::yield();
            else if (k == 63) std::this_thread::sleep_for(std::chrono::microseconds(rnd() % 200));
        };
        for (uint32_t seq = 1; seq <= (uint32_t)K && dead.load(std::memory_order_relaxed) == 0; ++seq) {
            pause();
            const size_t region = (size_t)(seq % (uint32_t)buffers) * (size_t)R;
            for (int i = 0; i < ll * R; ++i) {
                const int q = i / ll, w = i - q * ll;
                box[(size_t)q][(region + (size_t)me) * cap + w].store(((unsigned long long)seq << 32) | hh_mbox_payload(me, seq, w),
                                                                       std::memory_order_relaxed);
                if ((i & 7) == 0) pause();
            }
            for (int i = 0; i < ll * R; ++i) {
                const int q = i / ll, w = i - q * ll;
                const std::atomic<unsigned long long>& cell = box[(size_t)me][(region + (size_t)q) * cap + w];
                unsigned long long v = cell.load(std::memory_order_relaxed);
                int64_t spins = 0;
                while ((uint32_t)(v >> 32) != seq) {
                    if (++spins > spin_limit || dead.load(std::memory_order_relaxed) != 0) {
                        dead.store(1, std::memory_order_relaxed);
                        return;
                    }
                    if ((spins & 63) == 0) std::this_thread::yield();
                    v = cell.load(std::memory_order_relaxed);
                }
                if ((uint32_t)v != hh_mbox_payload(q, seq, w)) wrong.fetch_add(1, std::memory_order_relaxed);
                if (me == slow_rank) std::this_thread::sleep_for(std::chrono::microseconds(20));   // a rank that reads slowly
            }
        }
    };
    std::vector<std::thread> th;
    for (int r = 0; r < R; ++r) th.emplace_back(rank_fn, r);
    for (auto& t : th) t.join();
    return dead.load() != 0 ? -1 : wrong.load();
}
}  // extern "C"
