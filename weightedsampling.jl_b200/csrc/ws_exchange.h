// ws_exchange.h — the exchange plan of a sharded resampling step (pure host arithmetic, shared by the runtime and the
// CPU tests: tests/host/harness.cpp, tests/test_sharding_gloo.py).
//
// After the search every rank q has produced the ancestors of a contiguous range [bnd[2q], bnd[2q+1]) of GLOBAL output
// slots (the source rank searches the slots that fall into its CDF segment); the ranges are consecutive and
// cover [0, N).  Rank d owns the slots [lo(d), lo(d+1)), lo(d) = N d / R.  Everybody knows all bounds, so every rank
// derives the whole plan locally: what it sends where, what it receives from whom, where the pieces land.  Slot order
// makes every (source, destination) piece one contiguous range on both sides.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <deque>
#include <vector>

inline int64_t ws_rank_lo(int64_t n_global, int R, int d) { return (n_global * (int64_t)d) / R; }
// spare rows behind a shard's front planes (received offspring of the deferred gather land there)
inline int64_t ws_spare_rows(int64_t n_local) { return std::max<int64_t>(4096, n_local / 32); }
// slots of rank d's range that rank q produced
inline int64_t ws_piece(const int32_t* bnd, int64_t n_global, int R, int q, int d) {
    const int64_t lo = ws_rank_lo(n_global, R, d), hi = ws_rank_lo(n_global, R, d + 1);
    return std::max<int64_t>(0, std::min<int64_t>(bnd[2 * q + 1], hi) - std::max<int64_t>(bnd[2 * q], lo));
}

struct WsExchangePlan {
    int64_t fs = 0, fe = 0;                  // my produced slots [fs, fe)
    // what I send to d: produced slots in d's range = [send_off[d], send_off[d] + send_cnt[d]) of my produced range
    // what I get from q: q's produced slots in my range = my local slots [recv_off[q], recv_off[q] + recv_cnt[q])
    std::vector<int64_t> send_off, send_cnt, recv_off, recv_cnt;
    std::vector<int64_t> spare_pos;          // deferred gather: first spare row of rank q's offspring on MY side
    int64_t remote_send = 0, remote_recv = 0;
    int64_t total_remote = 0;                // migrants over all ranks (the same number on every rank)
    bool fits = true;                        // every rank's incoming offspring fit its spare rows (same on every rank)
};

inline WsExchangePlan ws_exchange_plan(const int32_t* bnd, int R, int r, int64_t n_global) {
    WsExchangePlan p;
    p.fs = bnd[2 * r];
    p.fe = bnd[2 * r + 1];
    p.send_off.assign(R, 0);
    p.send_cnt.assign(R, 0);
    p.recv_off.assign(R, 0);
    p.recv_cnt.assign(R, 0);
    p.spare_pos.assign(R, 0);
    const int64_t my_lo = ws_rank_lo(n_global, R, r), my_hi = ws_rank_lo(n_global, R, r + 1);
    for (int d = 0; d < R; ++d) {
        const int64_t a = std::max(p.fs, ws_rank_lo(n_global, R, d)), e = std::min(p.fe, ws_rank_lo(n_global, R, d + 1));
        p.send_cnt[d] = std::max<int64_t>(0, e - a);
        p.send_off[d] = std::max<int64_t>(0, a - p.fs);
        const int64_t qa = std::max<int64_t>(bnd[2 * d], my_lo), qe = std::min<int64_t>(bnd[2 * d + 1], my_hi);
        p.recv_cnt[d] = std::max<int64_t>(0, qe - qa);
        p.recv_off[d] = std::max<int64_t>(0, qa - my_lo);
        if (d != r) {
            p.remote_send += p.send_cnt[d];
            p.remote_recv += p.recv_cnt[d];
        }
    }
    int64_t acc = 0;
    for (int q = 0; q < R; ++q) {
        if (q == r) continue;
        p.spare_pos[q] = acc;
        acc += p.recv_cnt[q];
    }
    for (int d = 0; d < R; ++d) {
        const int64_t n_d = ws_rank_lo(n_global, R, d + 1) - ws_rank_lo(n_global, R, d);
        const int64_t incoming = n_d - ws_piece(bnd, n_global, R, d, d);
        p.total_remote += incoming;
        if (incoming > ws_spare_rows(n_d)) p.fits = false;
    }
    return p;
}

// Direct exchange, sender side: the element of rank d's plane buffer at which rank r's piece starts —
// deferred gather: rank d's spare rows (behind its n_d particles), in source-rank order;
// eager: the piece's final slots of rank d's back buffer.
inline int64_t ws_push_offset(const int32_t* bnd, int R, int r, int d, int64_t n_global, bool lazy) {
    const int64_t lo_d = ws_rank_lo(n_global, R, d);
    if (!lazy) return std::max<int64_t>(bnd[2 * r], lo_d) - lo_d;
    int64_t acc = 0;
    for (int q = 0; q < r; ++q)
        if (q != d) acc += ws_piece(bnd, n_global, R, q, d);
    return (ws_rank_lo(n_global, R, d + 1) - lo_d) + acc;
}

// ---- spare rows over several resampling events (sharded genealogy) ---------------------------------------------------
// A plane that is not read does not follow the resampling events: it keeps its own particle order and the
// offspring it received at event e stay in the spare rows they were pushed to, where the ancestor vector of event e
// points at them.  Spare rows are therefore handed out event by event, as a ring: the region of event e is released
// when no plane is older than e any more.  Every rank keeps the rings of ALL ranks (the incoming counts of every rank
// follow from the all-gathered bounds), so that a sender knows where its piece lands without asking.
struct WsSpareRegion {
    int64_t event, start, cnt;
};
typedef std::deque<WsSpareRegion> WsSpareRing;
// first row of a free run of `cnt` spare rows (capacity `cap`), or -1
inline int64_t ws_spare_ring_peek(const WsSpareRing& ring, int64_t cap, int64_t cnt) {
    if (cnt > cap) return -1;
    int64_t tail = -1, head = -1;   // start of the oldest / end of the newest non-empty region
    for (const auto& g : ring) {
        if (g.cnt <= 0) continue;
        if (tail < 0) tail = g.start;
        head = g.start + g.cnt;
    }
    if (tail < 0) return 0;
    if (head > tail) {              // occupied [tail, head): free behind the head, or in front of the tail
        if (head + cnt <= cap) return head;
        return cnt <= tail ? 0 : -1;
    }
    return head + cnt <= tail ? head : -1;   // wrapped: free [head, tail)
}
inline void ws_spare_ring_release(WsSpareRing& ring, int64_t upto_event) {   // regions of events <= upto_event
    while (!ring.empty() && ring.front().event <= upto_event) ring.pop_front();
}
