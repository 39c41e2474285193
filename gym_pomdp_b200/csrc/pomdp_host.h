// pomdp_host.h -- host-only glue shared by pomdp_kernels.cu (the product) and
// tests/hostsim/ (test vehicle): argument validation, conversion of the public parameter
// structs (include/pomdp_b200.h) into the by-value kernel parameter PODs, and construction
// of the static per-config maps.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pomdp_b200.h"
#include "pomdp_core.h"

namespace pomdp {
namespace host {

inline char* err_buf() {
    static thread_local char buf[256] = "";
    return buf;
}
inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 256, fmt, ap);
    va_end(ap);
    return code;
}

// binomial(1, p) == [r < T],  T = ceil(p * 2^32)   (p * 2^32 is exact in a double)
inline uint64_t bern_T(double p) {
    if (!(p > 0.0)) return 0;
    if (p >= 1.0) return 1ull << 32;
    return (uint64_t)ceil(p * 4294967296.0);
}
// (u > p) == [r > G],  G = floor(p * 2^32)
inline uint64_t gt_G(double p) {
    if (!(p > 0.0)) return 0;
    if (p >= 1.0) return 1ull << 32;
    return (uint64_t)floor(p * 4294967296.0);
}

// ---- RockSample benchmark layouts (the constants of rock.py:43-64) -----------------
struct RockLayout {
    int board, k_a, k_b;   // rock.py:101: num_rocks must be one of the two 'size' members
    int sx, sy, n_listed;
    int8_t pos[16][2];
};
inline const RockLayout* rock_layout(int board) {
    static const RockLayout L[] = {
        {2, 2, 1, 0, 0, 1, {{1, 0}}},
        {4, 4, 3, 0, 0, 3, {{1, 0}, {3, 1}, {2, 3}}},
        {7, 7, 8, 0, 3, 8, {{2, 0}, {0, 1}, {3, 1}, {6, 3}, {2, 4}, {3, 4}, {5, 5}, {1, 6}}},
        {11, 11, 11, 0, 5, 11,
         {{0, 3}, {0, 7}, {1, 8}, {2, 4}, {3, 3}, {3, 8}, {4, 3}, {5, 8}, {6, 1}, {9, 3}, {9, 9}}},
        {15, 15, 15, 0, 5, 16,
         {{0, 7}, {0, 3}, {1, 2}, {1, 2}, {2, 6}, {3, 7}, {3, 2}, {4, 7}, {5, 2}, {6, 9}, {9, 7}, {9, 1},
          {11, 8}, {13, 10}, {14, 9}, {12, 2}}},
    };
    for (const RockLayout& l : L)
        if (l.board == board) return &l;
    return nullptr;
}

inline int rock_words(const PomdpRockParams* q) { return q->num_rocks <= 11 ? 1 : 2; }

inline int rock_rows(const PomdpRockParams* q) { return 16 * (q->board_size - 1) + q->board_size; }  // cells x | y << 4 with x, y < n
inline int rock_lut_stride(const PomdpRockParams* q) { return (5 + q->num_rocks) | 1; }              // odd row pitch (rock_lut_index)
inline int64_t rock_table_bytes(const PomdpRockParams* q) {   // what pomdp_rock_build_table fills and the TMA copy moves
    const int64_t b = (int64_t)ROCK_LUT_OFFSET +
                      8 * ((int64_t)ROCK_SPECIALS + (int64_t)rock_rows(q) * rock_lut_stride(q) + ROCK_LUT_SKEW_MAX);
    return (b + 15) & ~(int64_t)15;
}
inline int64_t rock_smem_bytes(const PomdpRockParams* q) {
    return (int64_t)ROCK_LUT_OFFSET + 8 * ((int64_t)ROCK_SPECIALS + (int64_t)256 * rock_lut_stride(q) + ROCK_LUT_SKEW_MAX + 1);
}
inline uint32_t float_bits(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }
inline RockRes rock_result(int reward, int flags, int obs, bool done) {
    RockRes r;
    r.x = float_bits((float)reward);
    r.y = (uint32_t)obs;
    r.z = (uint32_t)(flags | (done ? FLAG_DONE : 0));
    r.w = done ? 0x80000000u : 0u;
    return r;
}

// Fills the kernel params and (if tbl != nullptr) the static maps (layout: pomdp_core.h).
// Returns 0 or POMDP_E_BADARG.
inline int make_rock(const PomdpRockParams* q, RockDev* d, void* tbl) {
    if (!q) return fail(POMDP_E_BADARG, "rock: params is NULL");
    const RockLayout* L = rock_layout(q->board_size);
    if (!L) return fail(POMDP_E_BADARG, "rock: board_size %d is not a key of rock.config (2,4,7,11,15)", q->board_size);
    if (q->num_rocks != L->k_a && q->num_rocks != L->k_b)
        return fail(POMDP_E_BADARG, "rock: num_rocks %d not in config[%d]['size'] (%d,%d)", q->num_rocks,
                    q->board_size, L->k_a, L->k_b);
    if (q->num_rocks > L->n_listed)
        return fail(POMDP_E_BADARG, "rock: config[%d] lists only %d rocks (the reference fails in _get_init_state)",
                    q->board_size, L->n_listed);
    const int n = q->board_size, k = q->num_rocks, n_act = 5 + k;
    const bool stoch = q->stochastic != 0;
    const bool wide = k > 11;                    // 64-bit packed states
    const int penal = stoch ? 0 : -100;
    if (d) {
        memset(d, 0, sizeof(*d));
        d->n = n;
        d->k = k;
        d->stochastic = stoch ? 1 : 0;
        d->penal = penal;
        d->start = (uint32_t)(L->sx | (L->sy << 4));
        d->n_actions = (uint32_t)n_act;
        d->lut_stride = (uint32_t)rock_lut_stride(q);
        d->table_bytes = (uint32_t)rock_table_bytes(q);
        d->smem_bytes = (uint32_t)rock_smem_bytes(q);
        const uint64_t T = bern_T(q->p_move);
        d->gate_on = T != 0;
        d->gate_thr_m1 = T ? (uint32_t)(T - 1) : 0u;
    }
    if (tbl) {
        RockTableHdr* h = (RockTableHdr*)tbl;
        RockRes* rtab = (RockRes*)((char*)tbl + ROCK_RTAB_OFFSET);
        RockLut* lut = (RockLut*)((char*)tbl + ROCK_LUT_OFFSET);
        memset(tbl, 0, (size_t)rock_table_bytes(q));
        memset(h->grid, -1, sizeof(h->grid));
        memset(h->rock_pos, 0xFF, sizeof(h->rock_pos));
        for (int i = 0; i < L->n_listed; ++i) {   // every listed rock is written; later ids overwrite (rock.py:110-111)
            const int cell = L->pos[i][0] | (L->pos[i][1] << 4);
            h->grid[cell] = (int8_t)i;
            h->rock_pos[i] = (uint8_t)cell;
        }
        for (int b = 0; b < 32; ++b) {            // rock.py:273-291: [E, N, S, W, SAMPLE, grid[rock.pos] + 5 ...]
            static const uint8_t head[5] = {1, 0, 2, 3, 4};
            if (b < 5) h->legal_act[b] = head[b];
            else if (b - 5 < k) h->legal_act[b] = (uint8_t)(5 + h->grid[h->rock_pos[b - 5]]);
            else h->legal_act[b] = 0xFF;
        }
        for (int dd = 0; dd < 32; ++dd) {
            const double eff = (1 + pow(2, -(double)dd / 20)) * .5;      // rock.py:383-387
            h->thr_m1[dd] = (uint32_t)(bern_T(eff) - 1);
            h->eff[dd] = eff;
        }
        // ---- result rows: entry = row + 2 * code + truthful; code 0 collected/none, 1 good, 3 bad (2 never packed)
        const bool wall_done = !stoch;                                   // rock.py:193; commented out at rock.py:503
        for (int c = 0; c < 4; ++c)
            for (int t = 0; t < 2; ++t) {
                const int j = 2 * c + t;
                rtab[ROCK_ROW_ZERO + j] = rock_result(0, 0, 0, false);
                rtab[ROCK_ROW_EXIT + j] = rock_result(10, 0, 0, true);                           // rock.py:139-141
                rtab[ROCK_ROW_WALL + j] = rock_result(penal, 0, 0, wall_done);
                const int rs = c == 0 ? penal : (c == 1 ? 10 : -10);                             // rock.py:160-169
                rtab[ROCK_ROW_SAMPLE + j] = rock_result(rs, 0, 0, c == 0 && wall_done);
                rtab[ROCK_ROW_DANGLING + j] = rock_result(penal, FLAG_BAD_STATE, 0, wall_done);  // IndexError at rock.py:162
                rtab[ROCK_ROW_CHECK + j] = rock_result(0, 0, ((c == 1) == (t == 1)) ? 2 : 1, false);  // rock.py:401-407
                rtab[ROCK_ROW_STEPPED_DONE + j] = rock_result(0, FLAG_DONE | FLAG_STEPPED_DONE, 0, false);
                rtab[ROCK_ROW_BAD_ACTION + j] = rock_result(0, FLAG_BAD_ACTION, 0, false);
            }
        const int none_bit = wide ? (int)RockBits<uint64_t>::NONE_SH1 + 1 : (int)RockBits<uint32_t>::NONE_SH1 + 1;
        // status_bit: bit offset of the rock status the action reads (or -1); clear: wipe those two bits; delta: cell xor
        auto entry = [&](uint32_t thr, int status_bit, bool clear, uint32_t delta, uint32_t row) {
            RockLut e;
            e.x = thr;
            e.y = (uint32_t)((status_bit < 0 ? none_bit : status_bit) - 1) | (row << 8) | (delta << 16) | ((clear ? 6u : 0u) << 24);
            return e;
        };
        lut[ROCK_IDX_NOOP] = entry(0xFFFFFFFFu, -1, false, 0, ROCK_ROW_ZERO);
        lut[ROCK_IDX_STEPPED_DONE] = entry(0xFFFFFFFFu, -1, false, 0, ROCK_ROW_STEPPED_DONE);
        lut[ROCK_IDX_BAD_ACTION] = entry(0xFFFFFFFFu, -1, false, 0, ROCK_ROW_BAD_ACTION);
        const int rows = rock_rows(q);
        for (int cell = 0; cell < rows; ++cell) {
            const int x = cell & 15, y = cell >> 4;
            RockLut* row = lut + ROCK_SPECIALS + (size_t)cell * rock_lut_stride(q) + (cell >> 4);        // rock_lut_index(cell, 0)
            for (int a = 0; a < 4; ++a) {                                // rock.py:134-158
                const int nx = x + move_dx(a), ny = y + move_dy(a);
                if ((unsigned)nx < (unsigned)n && (unsigned)ny < (unsigned)n)
                    row[a] = entry(0xFFFFFFFFu, -1, false, (uint32_t)(cell ^ (nx | (ny << 4))), ROCK_ROW_ZERO);
                else
                    row[a] = entry(0xFFFFFFFFu, -1, false, 0, a == 1 ? ROCK_ROW_EXIT : ROCK_ROW_WALL);
            }
            {                                                            // rock.py:160-169
                const int rock = h->grid[cell];
                if (rock >= k) row[4] = entry(0xFFFFFFFFu, -1, false, 0, ROCK_ROW_DANGLING);
                else if (rock >= 0) row[4] = entry(0xFFFFFFFFu, 8 + 2 * rock, true, 0, ROCK_ROW_SAMPLE);
                else row[4] = entry(0xFFFFFFFFu, -1, false, 0, ROCK_ROW_SAMPLE);
            }
            for (int r = 0; r < k; ++r) {                                // rock.py:171-175, 383-387, 401-407
                const int rp = h->rock_pos[r];
                row[5 + r] = entry(h->thr_m1[l1_distance(x, y, rp & 15, rp >> 4)], 8 + 2 * r, false, 0, ROCK_ROW_CHECK);
            }
        }
    }
    return 0;
}

inline int make_tag(const PomdpTagParams* q, TagDev* d) {
    if (!q) return fail(POMDP_E_BADARG, "tag: params is NULL");
    if (q->num_opponents < 1 || q->num_opponents > 4)
        return fail(POMDP_E_BADARG, "tag: num_opponents %d outside 1..4", q->num_opponents);
    memset(d, 0, sizeof(*d));
    d->n_opp = q->num_opponents;
    d->move_T = bern_T(q->move_prob);
    d->move_on = d->move_T != 0;
    d->move_thr_m1 = d->move_T ? (uint32_t)(d->move_T - 1) : 0u;
    return 0;
}

inline int make_tiger(const PomdpTigerParams* q, TigerDev* d) {
    if (!q) return fail(POMDP_E_BADARG, "tiger: params is NULL");
    d->listen_G = gt_G(q->listen_prob);
    d->listen_prob = q->listen_prob;
    return 0;
}

// The alias table of the joint failure draw (include/pomdp_b200.h, "Network draws"), in integer arithmetic only so that
// every implementation of the contract (this one, oracle/philox.py, oracle/pomdp_oracle.c) produces the same 256 columns.
//   digit counts  c0 = lo, c1 = hi - lo, c2 = 2^32 - hi          (lo = min(T_p, T_q), hi = max: u < lo / lo <= u < hi / else)
//   outcome k = sum d_i 3^i (d_i = digit of machine 5g + i), weight W[k] = (((c[d0] * c[d1] >> 32) * c[d2] >> 32) ...)
//   Vose over 256 columns with V[k] = 256 W[k] (0 for k >= 243) against the column mean S = sum W: small columns
//   (V < S, ascending k) and large ones are paired from the END of their lists; thr24 = floor(V * 2^24 / S).
inline void network_alias_table(uint64_t T_p, uint64_t T_q, NetAlias out[NET_COLS]) {
    const uint64_t lo = T_p < T_q ? T_p : T_q, hi = T_p < T_q ? T_q : T_p;
    const uint64_t c[3] = {lo, hi - lo, (1ull << 32) - hi};
    uint64_t V[NET_COLS] = {0}, S = 0;
    uint32_t own[NET_COLS] = {0};
    for (int k = 0; k < NET_CODES; ++k) {
        uint64_t a = 1ull << 32;
        uint32_t m_lo = 0, m_hi = 0;
        for (int i = 0, r = k; i < NET_GROUP; ++i, r /= 3) {
            const int d = r % 3;
            a = c[d] == (1ull << 32) ? a : (a * c[d]) >> 32;      // a <= 2^32 and c[d] < 2^32 here: no overflow
            if (d == 0) m_lo |= 1u << i;
            if (d <= 1) m_hi |= 1u << i;
        }
        V[k] = a;
        S += a;
        own[k] = m_lo | (m_hi << NET_GROUP);
    }
    int small[NET_COLS], large[NET_COLS], ns = 0, nl = 0, alias[NET_COLS];
    uint32_t thr24[NET_COLS];
    for (int k = 0; k < NET_COLS; ++k) {
        V[k] *= NET_COLS;
        alias[k] = k;
        thr24[k] = 0;                                              // a column that keeps its own outcome: alias = itself
        if (V[k] < S) small[ns++] = k; else large[nl++] = k;
    }
    while (ns > 0 && nl > 0) {
        const int sidx = small[--ns], lidx = large[--nl];
        thr24[sidx] = (uint32_t)((V[sidx] << 24) / S);              // V < S <= 2^32: fits
        alias[sidx] = lidx;
        V[lidx] -= S - V[sidx];
        if (V[lidx] < S) small[ns++] = lidx; else large[nl++] = lidx;
    }
    for (int k = 0; k < NET_COLS; ++k) {
        out[k].thr = thr24[k] << 8;
        out[k].masks = own[k] | (own[alias[k]] << 16);
    }
}

inline int make_network_uncached(const PomdpNetworkParams* q, NetworkDev* d);
// The alias columns and the neighbour-down map take ~10 us of host time to build -- half a kernel: the block of the last
// parameters seen is kept per thread (every launch of one env passes the same ones).
inline int make_network(const PomdpNetworkParams* q, NetworkDev* d) {
    if (!q) return fail(POMDP_E_BADARG, "network: params is NULL");
    struct Cache { bool valid; PomdpNetworkParams key; NetworkDev dev; };
    static thread_local Cache cache = {false, {}, {}};
    if (cache.valid && cache.key.n_machines == q->n_machines && cache.key.problem_type == q->problem_type &&
        memcmp(&cache.key.p, &q->p, sizeof(double)) == 0 && memcmp(&cache.key.q, &q->q, sizeof(double)) == 0 &&
        memcmp(&cache.key.p_ob, &q->p_ob, sizeof(double)) == 0) {
        *d = cache.dev;
        return 0;
    }
    const int rc = make_network_uncached(q, d);
    if (rc == 0) { cache.key = *q; cache.dev = *d; cache.valid = true; }
    return rc;
}
inline int make_network_uncached(const PomdpNetworkParams* q, NetworkDev* d) {
    if (!q) return fail(POMDP_E_BADARG, "network: params is NULL");
    const int n = q->n_machines;
    if (n < 1 || n > NETWORK_MAX) return fail(POMDP_E_BADARG, "network: n_machines %d outside 1..%d", n, NETWORK_MAX);
    memset(d, 0, sizeof(*d));
    d->n = n;
    d->p_T = bern_T(q->p);
    d->q_T = bern_T(q->q);
    d->ob_T = bern_T(q->p_ob);
    d->p_ob = q->p_ob;
    d->ob_any = d->ob_T != 0;
    d->om1 = (uint32_t)(d->ob_T - 1);
    int deg[NETWORK_MAX] = {0};
    auto link = [&](int i, int j) { d->nb[i] |= 1u << j; ++deg[i]; };
    if (q->problem_type == 3) {                     // network.py:153-168
        if (n < 4 || n % 3 != 1) return fail(POMDP_E_BADARG, "network: 3-legs needs n >= 4 and n %% 3 == 1 (network.py:155)");
        link(0, 1); link(0, 2); link(0, 3);
        for (int i = 1; i < n; ++i) {
            if (i < n - 3) link(i, i + 3);
            if (i <= 4) link(i, 0); else link(i, i - 3);
        }
    } else {                                        // network.py:144-151
        for (int i = 0; i < n; ++i) { link(i, (i + 1) % n); link(i, (i + n - 1) % n); }
    }
    for (int i = 0; i < n; ++i)
        if (deg[i] > 2) d->deg3 |= 1u << i;         // len(neighbours) counts duplicates too (network.py:89)
    d->groups = (n + NET_GROUP - 1) / NET_GROUP;
    d->cond_flip = d->q_T < d->p_T ? 0xFFFFFFFFu : 0u;
    // nbd[g][v]: the machines with a neighbour among the DOWN machines {5g + i : bit i of v} (network.py:81-84 is an OR
    // over neighbours, hence OR-linear in the down set: one lookup per group of five machines instead of one test per machine)
    for (int g = 0; g < d->groups; ++g)
        for (uint32_t v = 0; v < 32; ++v) {
            const uint32_t down = v << (NET_GROUP * g);
            uint32_t m = 0;
            for (int i = 0; i < n; ++i)
                if (d->nb[i] & down) m |= 1u << i;
            d->t.nbd[g][v] = m;
        }
    network_alias_table(d->p_T, d->q_T, d->t.alias);
    return 0;
}

inline int make_ship(const PomdpBattleshipParams* q, ShipDev* d) {
    if (!q) return fail(POMDP_E_BADARG, "battleship: params is NULL");
    if (q->x_size < 1 || q->y_size < 1 || q->x_size * q->y_size > SHIP_MAX_CELLS)
        return fail(POMDP_E_BADARG, "battleship: board %dx%d outside 1..%d cells", q->x_size, q->y_size, SHIP_MAX_CELLS);
    if (q->max_len < 2) return fail(POMDP_E_BADARG, "battleship: max_len %d < 2", q->max_len);
    int total = 0;
    for (int l = 2; l <= q->max_len; ++l) total += l;
    if (total > 127) return fail(POMDP_E_BADARG, "battleship: total ship length %d does not fit 7 bits", total);
    memset(d, 0, sizeof(*d));
    d->X = q->x_size; d->Y = q->y_size; d->max_len = q->max_len; d->n_tiles = q->x_size * q->y_size;
    B128 c0 = b128(0, 0), cL = b128(0, 0);
    for (int y = 0; y < d->Y; ++y) { c0 = c0 | b128_bit(y * d->X); cL = cL | b128_bit(y * d->X + d->X - 1); }
    d->col0_lo = c0.lo; d->col0_hi = c0.hi;
    d->colL_lo = cL.lo; d->colL_hi = cL.hi;
    if (q->max_len - 1 <= SHIP_MAX_SHIPS) {
        int ship = 0;
        for (int length = q->max_len; length >= 2; --length, ++ship) {
            B128 vp = b128(0, 0);
            for (int i = 0; i < length && i * d->X < 128; ++i) vp = vp | b128_bit(i * d->X);
            d->vpat_lo[ship] = vp.lo; d->vpat_hi[ship] = vp.hi;
            for (int dir = 0; dir < 4; ++dir) {
                B128 m = b128(0, 0);
                for (int y = 0; y < d->Y; ++y)
                    for (int x = 0; x < d->X; ++x)
                        if (grid_is_inside(d->X, d->Y, x + (length + 1) * move_dx(dir), y + (length + 1) * move_dy(dir)))
                            m = m | b128_bit(y * d->X + x);
                d->inside_lo[dir][ship] = m.lo; d->inside_hi[dir][ship] = m.hi;
            }
        }
    }
    return 0;
}

// The placement tables of BattleShip reset (layout: pomdp_core.h, ShipTableHdr).  With tbl == nullptr only the header
// is computed.  Returns the table's size in bytes (a multiple of 16) or a negative error code.
inline int64_t make_ship_table_from(const ShipDev& d, ShipTableHdr* hdr_out, void* tbl) {
    auto list_of = [&](const B128 valid[4], uint16_t* out) {     // accepted candidates in increasing c = 4 * pos + dir
        int n = 0;
        for (int pos = 0; pos < d.n_tiles; ++pos)
            for (int dir = 0; dir < 4; ++dir)
                if (b128_test(valid[dir], pos)) { if (out) out[n] = (uint16_t)(4 * pos + dir); ++n; }
        return n;
    };
    B128 valid0[4];
    ship_valid_starts(d, ship_blocked_b(d, b128(0, 0)), 0, d.max_len, valid0);
    uint16_t first[4 * SHIP_MAX_CELLS];
    const int n0 = list_of(valid0, first);
    const bool two = d.max_len >= 3;
    int64_t n_second = 0;
    if (two)
        for (int k0 = 0; k0 < n0; ++k0) {
            B128 v1[4];
            ship_valid_starts(d, ship_blocked_b(d, ship_cells(d, 0, first[k0] >> 2, first[k0] & 3, d.max_len)), 1, d.max_len - 1, v1);
            n_second += list_of(v1, nullptr);
        }
    auto up16 = [](int64_t v) { return (v + 15) & ~(int64_t)15; };
    ShipTableHdr h;
    memset(&h, 0, sizeof(h));
    h.magic = SHIP_TABLE_MAGIC; h.n0 = (uint32_t)n0; h.n_tabled = two ? 2u : 1u;
    h.off_rec = (uint32_t)sizeof(ShipTableHdr);
    h.off_second = (uint32_t)up16(h.off_rec + (int64_t)sizeof(ShipRec) * n0);
    h.bytes = (uint32_t)up16(h.off_second + 2 * n_second);
    if (hdr_out) *hdr_out = h;
    if (!tbl) return (int64_t)h.bytes;
    memset(tbl, 0, h.bytes);
    memcpy(tbl, &h, sizeof(h));
    ShipRec* rec = (ShipRec*)((char*)tbl + h.off_rec);
    uint16_t* t_second = (uint16_t*)((char*)tbl + h.off_second);
    uint32_t off = 0;
    for (int k0 = 0; k0 < n0; ++k0) {
        rec[k0].c0 = first[k0]; rec[k0].n1 = 0; rec[k0].off1 = off;
        if (!two) continue;
        B128 v1[4];
        ship_valid_starts(d, ship_blocked_b(d, ship_cells(d, 0, first[k0] >> 2, first[k0] & 3, d.max_len)), 1, d.max_len - 1, v1);
        const int n1 = list_of(v1, t_second + off);
        rec[k0].n1 = (uint16_t)n1;
        off += (uint32_t)n1;
    }
    return (int64_t)h.bytes;
}

// make_ship + the table layout fields of ShipDev (the enumeration takes ~50 us, so the last configuration is cached
// per thread: reset is called with the same params over and over)
inline int make_ship_tabled(const PomdpBattleshipParams* q, ShipDev* d) {
    int rc = make_ship(q, d);
    if (rc) return rc;
    if (q->max_len - 1 > SHIP_MAX_SHIPS) return fail(POMDP_E_BADARG, "battleship: more than %d ships", SHIP_MAX_SHIPS);
    static thread_local PomdpBattleshipParams cached_q = {0, 0, 0, 0};
    static thread_local ShipTableHdr cached_h;
    if (cached_q.x_size != q->x_size || cached_q.y_size != q->y_size || cached_q.max_len != q->max_len) {
        make_ship_table_from(*d, &cached_h, nullptr);
        cached_q = *q;
    }
    d->tbl_n0 = cached_h.n0; d->tbl_n_tabled = cached_h.n_tabled; d->tbl_off_rec = cached_h.off_rec; d->tbl_off_second = cached_h.off_second;
    return 0;
}
inline int64_t make_ship_table(const PomdpBattleshipParams* q, void* tbl) {
    ShipDev d;
    const int rc = make_ship(q, &d);
    if (rc) return rc < 0 ? rc : -rc;
    if (q->max_len - 1 > SHIP_MAX_SHIPS) return fail(POMDP_E_BADARG, "battleship: more than %d ships", SHIP_MAX_SHIPS);
    return make_ship_table_from(d, nullptr, tbl);
}

// pomdp_step_packed_host: the pipe, the kind, the state width the pipe was made for and the host pointers
inline int check_host_step(bool pipe_ok, int pipe_words, int kind, const void* params, const void* h_state, const void* h_action,
                           const void* h_next, const void* h_result, int64_t n, int64_t goff) {
    const char* what = "pomdp_step_packed_host";
    if (!pipe_ok) return fail(POMDP_E_BADARG, "%s: not a pipe (pomdp_host_pipe_create)", what);
    if (!params) return fail(POMDP_E_BADARG, "%s: params is NULL", what);
    int words = 1;
    if (kind == POMDP_KIND_ROCK) {
        const int rc = make_rock((const PomdpRockParams*)params, nullptr, nullptr);
        if (rc) return rc;
        words = rock_words((const PomdpRockParams*)params);
    } else if (kind != POMDP_KIND_TAG && kind != POMDP_KIND_TIGER && kind != POMDP_KIND_NETWORK) {
        return fail(POMDP_E_BADARG, "%s: kind %d has no packed step (Rock, Tag, Tiger, Network do)", what, kind);
    }
    if (words != pipe_words) return fail(POMDP_E_BADARG, "%s: the pipe was created for %d state words, this env has %d", what, pipe_words, words);
    if (n < 0 || goff < 0) return fail(POMDP_E_BADARG, "%s: n and global_offset must be >= 0", what);
    if (n > 0 && (!h_state || !h_action || !h_next || !h_result)) return fail(POMDP_E_BADARG, "%s: NULL host buffer", what);
    if ((((uintptr_t)h_state | (uintptr_t)h_action | (uintptr_t)h_next | (uintptr_t)h_result) & 3) != 0)
        return fail(POMDP_E_ALIGN, "%s: host buffers must be 4-byte aligned", what);
    return 0;
}

inline int hist_bins(int kind, int p0, int p1) {
    switch (kind) {
        case POMDP_KIND_ROCK: return p0 + 256;
        case POMDP_KIND_TAG: return 2 * TAG_CELLS;
        case POMDP_KIND_BATTLESHIP: return p0;
        case POMDP_KIND_TIGER: return 2;
        case POMDP_KIND_NETWORK: return p0;
    }
    (void)p1;
    return -1;
}

// shared argument checks of pomdp_coord_op / pomdp_belief_hist (host side, before any launch)
inline int check_coord_op(int op, int xs, const void* a, const void* b, const void* out, int64_t n) {
    if (op < 0 || op > POMDP_COORD_TAG_IS_INSIDE) return fail(POMDP_E_BADARG, "coord: unknown op %d", op);
    if (n < 0 || (n > 0 && (!a || !out))) return fail(POMDP_E_BADARG, "coord: bad n or NULL pointer");
    if ((op == POMDP_COORD_ADD_MOVE || op == POMDP_COORD_L1) && n > 0 && !b)
        return fail(POMDP_E_BADARG, "coord: op %d needs b", op);
    if ((op == POMDP_COORD_GET_COORD || op == POMDP_COORD_GET_INDEX) && xs <= 0)
        return fail(POMDP_E_BADARG, "coord: x_size must be positive");
    return 0;
}
inline int check_hist(int kind, int p0, int p1, const void* state, int words, int64_t n, const void* hist, int max_bins) {
    const int bins = hist_bins(kind, p0, p1);
    if (bins <= 0 || bins > max_bins) return fail(POMDP_E_BADARG, "belief_hist: bad kind/bins");
    if (words < 1 || words > SHIP_WORDS) return fail(POMDP_E_BADARG, "belief_hist: words %d outside 1..8", words);
    if (n < 0 || (n > 0 && (!state || !hist))) return fail(POMDP_E_BADARG, "belief_hist: bad n or NULL pointer");
    return 0;
}

inline int check_policy(const void* state, const void* action, int64_t n, int64_t goff, const char* what) {
    if (n < 0) return fail(POMDP_E_BADARG, "%s: n = %lld is negative", what, (long long)n);
    if (goff < 0) return fail(POMDP_E_BADARG, "%s: global_offset is negative", what);
    if (n == 0) return 0;
    if (!state || !action) return fail(POMDP_E_BADARG, "%s: a required array pointer is NULL", what);
    if (((uintptr_t)state | (uintptr_t)action) & 3) return fail(POMDP_E_ALIGN, "%s: array pointers must be 4-byte aligned", what);
    return 0;
}
inline int check_obs_prob(const void* state, const void* action, const void* obs, const void* prob, int64_t n, const char* what) {
    if (n < 0) return fail(POMDP_E_BADARG, "%s: n = %lld is negative", what, (long long)n);
    if (n == 0) return 0;
    if (!state || !action || !obs || !prob) return fail(POMDP_E_BADARG, "%s: a required array pointer is NULL", what);
    if ((((uintptr_t)state | (uintptr_t)action | (uintptr_t)obs) & 3) || ((uintptr_t)prob & 7))
        return fail(POMDP_E_ALIGN, "%s: int32 arrays must be 4-byte and the float64 array 8-byte aligned", what);
    return 0;
}
inline int check_belief(const void* state, const void* action, const void* obs, const void* count, const void* measured,
                        const void* lkv, const void* lkw, const void* pv, int64_t n) {
    if (n < 0) return fail(POMDP_E_BADARG, "pomdp_rock_belief_update: n = %lld is negative", (long long)n);
    if (n == 0) return 0;
    if (!state || !action || !obs || !count || !measured || !lkv || !lkw || !pv)
        return fail(POMDP_E_BADARG, "pomdp_rock_belief_update: a required array pointer is NULL");
    if ((((uintptr_t)state | (uintptr_t)action | (uintptr_t)obs | (uintptr_t)count | (uintptr_t)measured) & 3) ||
        (((uintptr_t)lkv | (uintptr_t)lkw | (uintptr_t)pv) & 7))
        return fail(POMDP_E_ALIGN, "pomdp_rock_belief_update: int32 arrays must be 4-byte and float64 arrays 8-byte aligned");
    return 0;
}
inline int check_rollout(const void* state, const void* final_state, const void* ret, const void* steps, const void* flags,
                         int64_t n, int64_t goff, int32_t max_steps, const char* what) {
    if (n < 0) return fail(POMDP_E_BADARG, "%s: n = %lld is negative", what, (long long)n);
    if (goff < 0) return fail(POMDP_E_BADARG, "%s: global_offset is negative", what);
    if (max_steps < 0) return fail(POMDP_E_BADARG, "%s: max_steps is negative", what);
    if (n == 0) return 0;
    if (!state || !ret || !steps || !flags) return fail(POMDP_E_BADARG, "%s: a required array pointer is NULL", what);
    if ((((uintptr_t)state | (uintptr_t)final_state | (uintptr_t)steps | (uintptr_t)flags) & 3) || ((uintptr_t)ret & 7))
        return fail(POMDP_E_ALIGN, "%s: int32 arrays must be 4-byte and the float64 return array 8-byte aligned", what);
    return 0;
}

inline int check_io(const void* state, const void* action, const void* next, const void* obs, const void* rw,
                    const void* fl, int64_t n) {
    if (n < 0) return fail(POMDP_E_BADARG, "n = %lld is negative", (long long)n);
    if (n == 0) return 0;
    if (!state || !action || !next || !obs || !rw || !fl) return fail(POMDP_E_BADARG, "a required array pointer is NULL");
    const uintptr_t any = (uintptr_t)state | (uintptr_t)action | (uintptr_t)next | (uintptr_t)obs | (uintptr_t)rw | (uintptr_t)fl;
    if (any & 3) return fail(POMDP_E_ALIGN, "array pointers must be 4-byte aligned");
    return 0;
}

}  // namespace host
}  // namespace pomdp
