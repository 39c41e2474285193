// pomdp_hostsim.cpp -- TEST VEHICLE, NOT A PRODUCT PATH.
//
// Compiles the per-env functors of gym_pomdp_b200/csrc/pomdp_core.h (the exact code the
// CUDA kernels inline) with g++ and exports the C ABI of include/pomdp_b200.h over HOST
// pointers, so that the packed-state logic can be checked against oracle/ in a container
// that has no GPU (SURVEY.md §4 tier 5).  Only tests/ loads the resulting
// tests/hostsim/libpomdp_hostsim.so; gym_pomdp_b200/ never does and has no CPU fallback.
// What this cannot cover -- and what the `-m gpu` tests are for -- is everything in
// pomdp_kernels.cu: the vector/TMA data movement, the warp-ballot placement scan, the
// shared-memory histogram, launch geometry.
#include <stdint.h>
#include <string.h>

#include "../../gym_pomdp_b200/csrc/pomdp_core.h"
#include "../../gym_pomdp_b200/csrc/pomdp_envs.h"
#include "../../gym_pomdp_b200/csrc/pomdp_host.h"

using namespace pomdp;

namespace {
template <typename S> S load_state(const int32_t* base, int64_t i);
template <> uint32_t load_state<uint32_t>(const int32_t* base, int64_t i) { return (uint32_t)base[i]; }
template <> uint64_t load_state<uint64_t>(const int32_t* base, int64_t i) {
    return (uint32_t)base[2 * i] | ((uint64_t)(uint32_t)base[2 * i + 1] << 32);
}
void store_state(int32_t* base, int64_t i, uint32_t s) { base[i] = (int32_t)s; }
void store_state(int32_t* base, int64_t i, uint64_t s) { base[2 * i] = (int32_t)(uint32_t)s; base[2 * i + 1] = (int32_t)(s >> 32); }

template <typename S>
int rock_step_host(const RockDev& d, const void* table, const int32_t* state, const int32_t* action, int32_t* next,
                   int32_t* obs, float* rw, int32_t* fl, int64_t n, int64_t goff, uint64_t seed, uint32_t step) {
    const RockRes* rtab = (const RockRes*)((const char*)table + ROCK_RTAB_OFFSET);
    const RockLut* lut = (const RockLut*)((const char*)table + ROCK_LUT_OFFSET);
    const PhiloxKey key = philox_key(seed);
    for (int64_t i = 0; i < n; ++i) {
        S s2;
        const uint64_t env = (uint64_t)(goff + i);
        const uint32_t wg = d.stochastic ? draw_word(key, env, step, DOMAIN_STEP, 0) : 0u;
        const uint32_t ws = draw_word(key, env, step, DOMAIN_STEP, 1);
        if (d.stochastic) rock_step<S, true>(d, lut, rtab, load_state<S>(state, i), action[i], wg, ws, s2, obs[i], rw[i], fl[i]);
        else rock_step<S, false>(d, lut, rtab, load_state<S>(state, i), action[i], wg, ws, s2, obs[i], rw[i], fl[i]);
        store_state(next, i, s2);
    }
    return 0;
}
}  // namespace

// ---- uniform-legal policy and fused rollouts: the SAME adapters (pomdp_envs.h) the kernels instantiate
namespace {
template <typename S> S load_any(const int32_t* base, int64_t i) { return load_state<S>(base, i); }

template <class Env>
int policy_host(const typename Env::Params& d, const void* table, const int32_t* state, int32_t* action, int64_t n,
                int64_t goff, uint64_t seed, uint32_t step, const char* what) {
    int rc = host::check_policy(state, action, n, goff, what);
    if (rc) return rc;
    const PhiloxKey key = philox_key(seed);
    const unsigned char* tbl = (const unsigned char*)table;
    for (int64_t i = 0; i < n; ++i)
        action[i] = Env::policy(d, tbl, load_any<typename Env::State>(state, i),
                                draw_word(key, (uint64_t)(goff + i), step, DOMAIN_POLICY, 0));
    return 0;
}
template <class Env>
int rollout_host(const typename Env::Params& d, const void* table, const int32_t* state, const int32_t* first_action,
                 int32_t* final_state, double* ret,
                 int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step, int32_t max_steps,
                 double gamma, const char* what) {
    int rc = host::check_rollout(state, final_state, ret, steps, flags, n, goff, max_steps, what);
    if (rc) return rc;
    const PhiloxKey key = philox_key(seed);
    const unsigned char* tbl = (const unsigned char*)table;
    for (int64_t i = 0; i < n; ++i) {
        typename Env::State s = load_any<typename Env::State>(state, i);
        RolloutAcc acc;
        rollout1<Env>(d, tbl, s, key, (uint64_t)(goff + i), step, max_steps, gamma, acc, first_action != nullptr,
                      first_action ? first_action[i] : 0);
        if (final_state) store_state(final_state, i, s);
        ret[i] = acc.ret; steps[i] = acc.steps; flags[i] = acc.flags;
    }
    return 0;
}
}  // namespace

namespace {
template <class Env>
int obs_prob_host(const typename Env::Params& d, const void* table, const int32_t* state, const int32_t* action, const int32_t* obs,
                  double* prob, int64_t n, double extra, const char* what) {
    int rc = host::check_obs_prob(state, action, obs, prob, n, what);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i)
        prob[i] = Env::obs_prob(d, (const unsigned char*)table, load_any<typename Env::State>(state, i), action[i], obs[i], extra);
    return 0;
}
template <class Env>
int legal_mask_host(const typename Env::Params& d, const void* table, const int32_t* state, uint32_t* mask, int64_t n,
                    const char* what) {
    int rc = host::check_policy(state, mask, n, 0, what);
    if (rc) return rc;
    const int words = Env::mask_words(d);
    for (int64_t i = 0; i < n; ++i) {
        uint32_t m[2] = {0u, 0u};
        Env::legal_mask(d, (const unsigned char*)table, load_any<typename Env::State>(state, i), m);
        for (int k = 0; k < words; ++k) mask[i * words + k] = m[k];
    }
    return 0;
}
}  // namespace

namespace {
// *_step_packed on the host: the unpacked entry point, then pack_result per env (what the kernels' epilogue does)
template <class Env, class F>
int step_packed_host(F step, const int32_t* state, const int32_t* action, int32_t* next, int32_t* result, int64_t n) {
    if (n <= 0) return step(state, action, next, result, (float*)result, result, n);
    int32_t* ob = new int32_t[n];
    int32_t* fl = new int32_t[n];
    float* rw = new float[n];
    const int rc = step(state, action, next, ob, rw, fl, n);
    if (!rc)
        for (int64_t i = 0; i < n; ++i) result[i] = pack_result(ob[i], Env::reward_units(rw[i]), fl[i]);
    delete[] ob; delete[] fl; delete[] rw;
    return rc;
}
}  // namespace

extern "C" {

int pomdp_abi_version(void) { return POMDP_ABI_VERSION; }
const char* pomdp_last_error(void) { return host::err_buf(); }
int pomdp_is_hostsim(void) { return 1; }   // lets tests assert which library they are talking to

int pomdp_rock_state_words(const PomdpRockParams* q) {
    int rc = host::make_rock(q, nullptr, nullptr);
    return rc ? rc : host::rock_words(q);
}
int64_t pomdp_rock_table_bytes(const PomdpRockParams* q) {
    int rc = host::make_rock(q, nullptr, nullptr);
    return rc ? (int64_t)rc : host::rock_table_bytes(q);
}
int pomdp_rock_build_table(const PomdpRockParams* q, void* host_table) {
    if (!host_table) return host::fail(POMDP_E_BADARG, "rock: host_table is NULL");
    RockDev d;
    return host::make_rock(q, &d, host_table);
}
int pomdp_rock_step(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* action,
                    int32_t* next, int32_t* obs, float* rw, int32_t* fl, int64_t n, int64_t goff, uint64_t seed,
                    uint32_t step, void*) {
    RockDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    rc = host::check_io(state, action, next, obs, rw, fl, n);
    if (rc) return rc;
    if (!table) return host::fail(POMDP_E_BADARG, "rock: table is NULL");
    const void* t = table;
    return host::rock_words(q) == 1 ? rock_step_host<uint32_t>(d, t, state, action, next, obs, rw, fl, n, goff, seed, step)
                                    : rock_step_host<uint64_t>(d, t, state, action, next, obs, rw, fl, n, goff, seed, step);
}
int pomdp_rock_reset(const PomdpRockParams* q, const void*, int32_t* state, int32_t* obs, const uint8_t* mask, int64_t n,
                     int64_t goff, uint64_t seed, uint32_t step, void*) {
    RockDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        const LazyDraw draw{&key, (uint64_t)(goff + i), step, DOMAIN_RESET};
        if (host::rock_words(q) == 1) store_state(state, i, rock_reset<uint32_t>(d, draw));
        else store_state(state, i, rock_reset<uint64_t>(d, draw));
        if (obs) obs[i] = 0;
    }
    return 0;
}

int64_t pomdp_tag_table_bytes(void) { return (int64_t)sizeof(TagTables); }
int pomdp_tag_build_table(void* host_table) {
    if (!host_table) return host::fail(POMDP_E_BADARG, "tag: host_table is NULL");
    tag_build_tables((TagTables*)host_table);
    return 0;
}
int pomdp_tag_step(const PomdpTagParams* q, const void* table, const int32_t* state, const int32_t* action, int32_t* next,
                   int32_t* obs, float* rw, int32_t* fl, int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    TagDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    rc = host::check_io(state, action, next, obs, rw, fl, n);
    if (rc) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "tag: table is NULL");
    const TagTables* T = (const TagTables*)table;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t s2;
        const LazyDraw draw{&key, (uint64_t)(goff + i), step, DOMAIN_STEP};
        // the stock one-opponent env alternates between the kernels' two functors (general / branch-free)
        if (d.n_opp == 1 && (i & 1)) tag_step_1opp(d, T, (uint32_t)state[i], action[i], draw(0), s2, obs[i], rw[i], fl[i]);
        else tag_step(d, T, (uint32_t)state[i], action[i], draw, s2, obs[i], rw[i], fl[i]);
        next[i] = (int32_t)s2;
    }
    return 0;
}
int pomdp_tag_reset(const PomdpTagParams* q, int32_t* state, int32_t* obs, const uint8_t* mask, int64_t n, int64_t goff,
                    uint64_t seed, uint32_t step, void*) {
    TagDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        uint32_t s; int32_t ob;
        tag_reset(d, LazyDraw{&key, (uint64_t)(goff + i), step, DOMAIN_RESET}, s, ob);
        state[i] = (int32_t)s;
        if (obs) obs[i] = ob;
    }
    return 0;
}

int pomdp_tiger_step(const PomdpTigerParams* q, const int32_t* state, const int32_t* action, int32_t* next, int32_t* obs,
                     float* rw, int32_t* fl, int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    TigerDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    rc = host::check_io(state, action, next, obs, rw, fl, n);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t s2;
        tiger_step(d, (uint32_t)state[i], action[i], LazyDraw{&key, (uint64_t)(goff + i), step, DOMAIN_STEP}, s2, obs[i], rw[i], fl[i]);
        next[i] = (int32_t)s2;
    }
    return 0;
}
int pomdp_tiger_reset(const PomdpTigerParams* q, int32_t* state, int32_t* obs, const uint8_t* mask, int64_t n,
                      int64_t goff, uint64_t seed, uint32_t step, void*) {
    TigerDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        uint32_t s; int32_t ob;
        tiger_reset(LazyDraw{&key, (uint64_t)(goff + i), step, DOMAIN_RESET}, s, ob);
        state[i] = (int32_t)s;
        if (obs) obs[i] = ob;
    }
    return 0;
}

int pomdp_network_step(const PomdpNetworkParams* q, const int32_t* state, const int32_t* action, int32_t* next,
                       int32_t* obs, float* rw, int32_t* fl, int64_t n, int64_t goff, uint64_t seed, uint32_t step,
                       void*) {
    NetworkDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    rc = host::check_io(state, action, next, obs, rw, fl, n);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        // alternate the two flavours the kernels use: aligned groups of four, and single envs
        const uint64_t env = (uint64_t)(goff + i);
        if ((env & 3) == 0 && i + 4 <= n && ((i >> 2) & 1) == 0) {
            uint32_t s4[4], s2[4];
            for (int j = 0; j < 4; ++j) s4[j] = (uint32_t)state[i + j];
            if (d.groups == 2)      // the compile-time two-group flavour the kernels use for the stock 10-machine network
                network_step_n<4, NetTabPtr, 2>(d, NetTabPtr{&d.t}, s4, action + i, key, env >> 2, 0, step, s2, obs + i, rw + i, fl + i);
            else
                network_step_n<4>(d, NetTabPtr{&d.t}, s4, action + i, key, env >> 2, 0, step, s2, obs + i, rw + i, fl + i);
            for (int j = 0; j < 4; ++j) next[i + j] = (int32_t)s2[j];
            i += 3;
            continue;
        }
        uint32_t s1 = (uint32_t)state[i], s2;
        network_step_n<1>(d, NetTabPtr{&d.t}, &s1, action + i, key, env >> 2, (int)(env & 3), step, &s2, obs + i, rw + i, fl + i);
        next[i] = (int32_t)s2;
    }
    return 0;
}
int pomdp_network_reset(const PomdpNetworkParams* q, int32_t* state, int32_t* obs, const uint8_t* mask, int64_t n, void*) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        state[i] = (int32_t)((1u << d.n) - 1u);
        if (obs) obs[i] = 0;
    }
    return 0;
}

int pomdp_battleship_step(const PomdpBattleshipParams* q, const int32_t* state, const int32_t* action, int32_t* next,
                          int32_t* obs, float* rw, int32_t* fl, int64_t n, void*) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    rc = host::check_io(state, action, next, obs, rw, fl, n);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t w[SHIP_WORDS], w2[SHIP_WORDS];
        for (int k = 0; k < SHIP_WORDS; ++k) w[k] = (uint32_t)state[i * SHIP_WORDS + k];
        battleship_step(d, w, action[i], w2, obs[i], rw[i], fl[i]);
        for (int k = 0; k < SHIP_WORDS; ++k) next[i * SHIP_WORDS + k] = (int32_t)w2[k];
    }
    return 0;
}
int64_t pomdp_battleship_table_bytes(const PomdpBattleshipParams* q) { return host::make_ship_table(q, nullptr); }
int pomdp_battleship_build_table(const PomdpBattleshipParams* q, void* host_table) {
    if (!host_table) return host::fail(POMDP_E_BADARG, "battleship: host_table is NULL");
    const int64_t rc = host::make_ship_table(q, host_table);
    return rc < 0 ? (int)-rc : 0;
}
// table == NULL: the bitboard scan; otherwise the placement tables (the functors the two kernels inline)
int pomdp_battleship_reset(const PomdpBattleshipParams* q, const void* table, int32_t* state, int32_t* obs, int32_t* flags,
                           const uint8_t* mask, int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    ShipDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = table ? host::make_ship_tabled(q, &d) : host::make_ship(q, &d);
    if (rc) return rc;
    if (q->max_len - 1 > SHIP_MAX_SHIPS) return host::fail(POMDP_E_BADARG, "battleship: more than 8 ships");
    if ((uintptr_t)table & 15) return host::fail(POMDP_E_ALIGN, "pomdp_battleship_reset: d_table must be 16-byte aligned");
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        ShipState st;
        const bool lean = d.max_len <= 3 && d.tbl_n_tabled >= (uint32_t)(d.max_len - 1);
        const bool ok = table ? (lean ? battleship_reset_table<true>(d, (const unsigned char*)table, ShipDraw(key, (uint64_t)(goff + i), step), st)
                                      : battleship_reset_table<false>(d, (const unsigned char*)table, ShipDraw(key, (uint64_t)(goff + i), step), st))
                              : battleship_reset_bitboard(d, key, (uint64_t)(goff + i), step, st);
        uint32_t w8[SHIP_WORDS];
        ship_pack(st, w8);
        for (int k = 0; k < SHIP_WORDS; ++k) state[i * SHIP_WORDS + k] = (int32_t)w8[k];
        if (obs) obs[i] = 0;
        if (flags) flags[i] = ok ? 0 : FLAG_BAD_STATE;
    }
    return 0;
}
// Serial stand-in for the warp scan: same candidate order, same k-th-accepted rule.
int pomdp_battleship_reset_warpscan(const PomdpBattleshipParams* q, int32_t* state, int32_t* obs, int32_t* flags,
                           const uint8_t* mask, int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    ShipDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        ShipState st;
        st.occ = b128(0, 0); st.vis = b128(0, 0); st.remaining = 0; st.done = false;
        bool ok_all = true;
        int ship = 0;
        for (int length = d.max_len; length >= 2 && ok_all; --length, ++ship) {
            const B128 blocked = ship_blocked(d, st.occ);
            int total = 0;
            for (int c = 0; c < 4 * d.n_tiles; ++c) total += ship_candidate_ok(d, blocked, c >> 2, c & 3, length);
            if (total == 0) { ok_all = false; break; }
            int k = (int)rand_below(ShipDraw(key, (uint64_t)(goff + i), step)(ship), (uint32_t)total);
            for (int c = 0; c < 4 * d.n_tiles; ++c)
                if (ship_candidate_ok(d, blocked, c >> 2, c & 3, length) && k-- == 0) { ship_mark(d, st, c >> 2, c & 3, length); break; }
        }
        uint32_t w8[SHIP_WORDS];
        ship_pack(st, w8);
        for (int k = 0; k < SHIP_WORDS; ++k) state[i * SHIP_WORDS + k] = (int32_t)w8[k];
        if (obs) obs[i] = 0;
        if (flags) flags[i] = ok_all ? 0 : FLAG_BAD_STATE;
    }
    return 0;
}
int pomdp_battleship_reset_rejection(const PomdpBattleshipParams* q, int32_t* state, int32_t* obs, int32_t* flags,
                                     const uint8_t* mask, int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    ShipDev d;
    const PhiloxKey key = philox_key(seed);
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        ShipState st;
        const bool ok = battleship_reset_rejection(d, key, (uint64_t)(goff + i), step, st, 4096);
        uint32_t w8[SHIP_WORDS];
        ship_pack(st, w8);
        for (int k = 0; k < SHIP_WORDS; ++k) state[i * SHIP_WORDS + k] = (int32_t)w8[k];
        if (obs) obs[i] = 0;
        if (flags) flags[i] = ok ? 0 : FLAG_BAD_STATE;
    }
    return 0;
}

#define HOSTSIM_ROCK_DISPATCH(FN, ...)                                                        \
    do {                                                                                      \
        if (!table) return host::fail(POMDP_E_BADARG, "rock: table is NULL");                 \
        if (host::rock_words(q) == 1) {                                                       \
            if (d.stochastic) return FN<RockEnvT<uint32_t, true>>(__VA_ARGS__);               \
            return FN<RockEnvT<uint32_t, false>>(__VA_ARGS__);                                \
        }                                                                                     \
        if (d.stochastic) return FN<RockEnvT<uint64_t, true>>(__VA_ARGS__);                   \
        return FN<RockEnvT<uint64_t, false>>(__VA_ARGS__);                                    \
    } while (0)

int pomdp_rock_policy(const PomdpRockParams* q, const void* table, const int32_t* state, int32_t* action, int64_t n,
                      int64_t goff, uint64_t seed, uint32_t step, void*) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    HOSTSIM_ROCK_DISPATCH(policy_host, d, table, state, action, n, goff, seed, step, "pomdp_rock_policy");
}
int pomdp_rock_rollout(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret,
                       int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step,
                       int32_t max_steps, double gamma, void*) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    HOSTSIM_ROCK_DISPATCH(rollout_host, d, table, state, first_action, final_state, ret, steps, flags, n, goff, seed, step, max_steps, gamma,
                          "pomdp_rock_rollout");
}
int pomdp_tag_policy(const PomdpTagParams* q, const void* table, const int32_t* state, int32_t* action, int64_t n, int64_t goff,
                     uint64_t seed, uint32_t step, void*) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    return policy_host<TagEnvT<1>>(d, table, state, action, n, goff, seed, step, "pomdp_tag_policy");
}
int pomdp_tag_rollout(const PomdpTagParams* q, const void* table, const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret,
                      int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step, int32_t max_steps,
                      double gamma, void*) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "tag: table is NULL");
    if (d.n_opp == 1)
        return rollout_host<TagEnvT<1>>(d, table, state, first_action, final_state, ret, steps, flags, n, goff, seed, step, max_steps, gamma,
                                        "pomdp_tag_rollout");
    return rollout_host<TagEnvT<4>>(d, table, state, first_action, final_state, ret, steps, flags, n, goff, seed, step, max_steps, gamma,
                                    "pomdp_tag_rollout");
}
int pomdp_tiger_policy(const PomdpTigerParams* q, const int32_t* state, int32_t* action, int64_t n, int64_t goff,
                       uint64_t seed, uint32_t step, void*) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return policy_host<TigerEnvP>(d, nullptr, state, action, n, goff, seed, step, "pomdp_tiger_policy");
}
int pomdp_tiger_rollout(const PomdpTigerParams* q, const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret, int32_t* steps,
                        int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step, int32_t max_steps,
                        double gamma, void*) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return rollout_host<TigerEnvP>(d, nullptr, state, first_action, final_state, ret, steps, flags, n, goff, seed, step, max_steps, gamma,
                                   "pomdp_tiger_rollout");
}
int pomdp_network_policy(const PomdpNetworkParams* q, const int32_t* state, int32_t* action, int64_t n, int64_t goff,
                         uint64_t seed, uint32_t step, void*) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    return policy_host<NetworkEnvP>(d, nullptr, state, action, n, goff, seed, step, "pomdp_network_policy");
}
int pomdp_network_rollout(const PomdpNetworkParams* q, const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret,
                          int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step,
                          int32_t max_steps, double gamma, void*) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    return rollout_host<NetworkEnvP>(d, nullptr, state, first_action, final_state, ret, steps, flags, n, goff, seed, step, max_steps, gamma,
                                     "pomdp_network_rollout");
}
int pomdp_battleship_policy(const PomdpBattleshipParams* q, const int32_t* state, int32_t* action, int64_t n, int64_t goff,
                            uint64_t seed, uint32_t step, void*) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if ((rc = host::check_policy(state, action, n, goff, "pomdp_battleship_policy"))) return rc;
    const PhiloxKey key = philox_key(seed);
    for (int64_t i = 0; i < n; ++i)
        action[i] = battleship_policy(d, (const uint32_t*)state + i * SHIP_WORDS,
                                      draw_word(key, (uint64_t)(goff + i), step, DOMAIN_POLICY, 0));
    return 0;
}
int pomdp_battleship_rollout(const PomdpBattleshipParams* q, const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret,
                             int32_t* steps, int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step,
                             int32_t max_steps, double gamma, void*) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if ((rc = host::check_rollout(state, final_state, ret, steps, flags, n, goff, max_steps, "pomdp_battleship_rollout")))
        return rc;
    const PhiloxKey key = philox_key(seed);
    for (int64_t i = 0; i < n; ++i) {
        uint32_t w[SHIP_WORDS], w2[SHIP_WORDS];
        memcpy(w, state + i * SHIP_WORDS, sizeof(w));
        RolloutAcc acc;
        acc.init((w[3] >> 31) != 0);
        for (int32_t t = 0; t < max_steps && !(w[3] >> 31); ++t) {
            const int32_t a = (t == 0 && first_action)
                                  ? first_action[i]
                                  : battleship_policy(d, w, draw_word(key, (uint64_t)(goff + i), step + (uint32_t)t, DOMAIN_POLICY, 0));
            int32_t ob, fl;
            float rw;
            battleship_step(d, w, a, w2, ob, rw, fl);
            memcpy(w, w2, sizeof(w));
            acc.add((double)rw, gamma, fl);
        }
        if (final_state) memcpy(final_state + i * SHIP_WORDS, w, sizeof(w));
        ret[i] = acc.ret; steps[i] = acc.steps; flags[i] = acc.flags;
    }
    return 0;
}

int pomdp_rock_obs_prob(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* action, const int32_t* obs,
                        double* prob, int64_t n, void*) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "rock: table is NULL");
    if (host::rock_words(q) == 1) return obs_prob_host<RockEnvT<uint32_t, false>>(d, table, state, action, obs, prob, n, 0.0, "pomdp_rock_obs_prob");
    return obs_prob_host<RockEnvT<uint64_t, false>>(d, table, state, action, obs, prob, n, 0.0, "pomdp_rock_obs_prob");
}
int pomdp_rock_legal_mask(const PomdpRockParams* q, const void* table, const int32_t* state, uint32_t* mask, int64_t n, void*) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "rock: table is NULL");
    if (host::rock_words(q) == 1) return legal_mask_host<RockEnvT<uint32_t, false>>(d, table, state, mask, n, "pomdp_rock_legal_mask");
    return legal_mask_host<RockEnvT<uint64_t, false>>(d, table, state, mask, n, "pomdp_rock_legal_mask");
}
int pomdp_tag_obs_prob(const PomdpTagParams* q, const int32_t* state, const int32_t* action, const int32_t* obs, double* prob,
                       int64_t n, void*) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    return obs_prob_host<TagNoTable>(d, nullptr, state, action, obs, prob, n, 0.0, "pomdp_tag_obs_prob");
}
int pomdp_tag_legal_mask(const PomdpTagParams* q, const int32_t* state, uint32_t* mask, int64_t n, void*) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    return legal_mask_host<TagNoTable>(d, nullptr, state, mask, n, "pomdp_tag_legal_mask");
}
int pomdp_tiger_obs_prob(const PomdpTigerParams* q, const int32_t* state, const int32_t* action, const int32_t* obs, double* prob,
                         int64_t n, double correct_prob, void*) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return obs_prob_host<TigerEnvP>(d, nullptr, state, action, obs, prob, n, correct_prob, "pomdp_tiger_obs_prob");
}
int pomdp_tiger_legal_mask(const PomdpTigerParams* q, const int32_t* state, uint32_t* mask, int64_t n, void*) {
    TigerDev d;
    int rc = host::make_tiger(q, &d);
    if (rc) return rc;
    return legal_mask_host<TigerEnvP>(d, nullptr, state, mask, n, "pomdp_tiger_legal_mask");
}
int pomdp_network_obs_prob(const PomdpNetworkParams* q, const int32_t* state, const int32_t* action, const int32_t* obs,
                           double* prob, int64_t n, void*) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    return obs_prob_host<NetworkEnvP>(d, nullptr, state, action, obs, prob, n, 0.0, "pomdp_network_obs_prob");
}
int pomdp_network_legal_mask(const PomdpNetworkParams* q, const int32_t* state, uint32_t* mask, int64_t n, void*) {
    NetworkDev d;
    int rc = host::make_network(q, &d);
    if (rc) return rc;
    return legal_mask_host<NetworkEnvP>(d, nullptr, state, mask, n, "pomdp_network_legal_mask");
}
int pomdp_battleship_obs_prob(const PomdpBattleshipParams* q, const int32_t* state, const int32_t* action, const int32_t* obs,
                              double* prob, int64_t n, void*) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if ((rc = host::check_obs_prob(state, action, obs, prob, n, "pomdp_battleship_obs_prob"))) return rc;
    for (int64_t i = 0; i < n; ++i) prob[i] = battleship_obs_prob(d, (const uint32_t*)state + i * SHIP_WORDS, action[i], obs[i]);
    return 0;
}
int pomdp_battleship_legal_mask(const PomdpBattleshipParams* q, const int32_t* state, uint32_t* mask, int64_t n, void*) {
    ShipDev d;
    int rc = host::make_ship(q, &d);
    if (rc) return rc;
    if ((rc = host::check_policy(state, mask, n, 0, "pomdp_battleship_legal_mask"))) return rc;
    const int words = (d.n_tiles + 31) >> 5;
    for (int64_t i = 0; i < n; ++i)
        for (int k = 0; k < words; ++k) {
            const int lim = d.n_tiles - 32 * k;
            const uint32_t valid = lim >= 32 ? 0xFFFFFFFFu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
            mask[i * words + k] = ~(uint32_t)state[i * SHIP_WORDS + 4 + k] & valid;
        }
    return 0;
}

int pomdp_rock_step_packed(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* action, int32_t* next,
                           int32_t* result, int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    if (n > 0 && !result) return host::fail(POMDP_E_BADARG, "a required array pointer is NULL");
    return step_packed_host<RockEnvT<uint32_t, false>>(
        [&](const int32_t* s, const int32_t* a, int32_t* nx, int32_t* ob, float* rw, int32_t* fl, int64_t m) {
            return pomdp_rock_step(q, table, s, a, nx, ob, rw, fl, m, goff, seed, step, nullptr);
        }, state, action, next, result, n);
}
int pomdp_tag_step_packed(const PomdpTagParams* q, const void* table, const int32_t* state, const int32_t* action, int32_t* next,
                          int32_t* result, int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    if (n > 0 && !result) return host::fail(POMDP_E_BADARG, "a required array pointer is NULL");
    return step_packed_host<TagEnvT<1>>(
        [&](const int32_t* s, const int32_t* a, int32_t* nx, int32_t* ob, float* rw, int32_t* fl, int64_t m) {
            return pomdp_tag_step(q, table, s, a, nx, ob, rw, fl, m, goff, seed, step, nullptr);
        }, state, action, next, result, n);
}
int pomdp_tiger_step_packed(const PomdpTigerParams* q, const int32_t* state, const int32_t* action, int32_t* next, int32_t* result,
                            int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    if (n > 0 && !result) return host::fail(POMDP_E_BADARG, "a required array pointer is NULL");
    return step_packed_host<TigerEnvP>(
        [&](const int32_t* s, const int32_t* a, int32_t* nx, int32_t* ob, float* rw, int32_t* fl, int64_t m) {
            return pomdp_tiger_step(q, s, a, nx, ob, rw, fl, m, goff, seed, step, nullptr);
        }, state, action, next, result, n);
}
int pomdp_network_step_packed(const PomdpNetworkParams* q, const int32_t* state, const int32_t* action, int32_t* next,
                              int32_t* result, int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    if (n > 0 && !result) return host::fail(POMDP_E_BADARG, "a required array pointer is NULL");
    return step_packed_host<NetworkEnvP>(
        [&](const int32_t* s, const int32_t* a, int32_t* nx, int32_t* ob, float* rw, int32_t* fl, int64_t m) {
            return pomdp_network_step(q, s, a, nx, ob, rw, fl, m, goff, seed, step, nullptr);
        }, state, action, next, result, n);
}

// host-buffer pipeline: no device, so a "pipe" only remembers its chunking; the chunks go through the packed steps
// one after the other with the same global offsets the CUDA library uses.
struct HostPipe { uint32_t magic; int words, n_slots; int64_t chunk; };
int pomdp_host_pipe_create(int state_words, int64_t chunk_envs, int n_slots, void** pipe_out) {
    if (!pipe_out) return host::fail(POMDP_E_BADARG, "pomdp_host_pipe_create: pipe_out is NULL");
    *pipe_out = nullptr;
    if (state_words < 1 || state_words > SHIP_WORDS || chunk_envs < 4 || (chunk_envs & 3) || chunk_envs > (1ll << 28) ||
        n_slots < 1 || n_slots > 8)
        return host::fail(POMDP_E_BADARG, "pomdp_host_pipe_create: state_words %d, chunk_envs %lld (multiple of 4), n_slots %d (1..8)",
                          state_words, (long long)chunk_envs, n_slots);
    *pipe_out = new HostPipe{0x50495045u, state_words, n_slots, chunk_envs};
    return 0;
}
int pomdp_host_pipe_destroy(void* pipe) {
    HostPipe* hp = (HostPipe*)pipe;
    if (!hp || hp->magic != 0x50495045u) return host::fail(POMDP_E_BADARG, "pomdp_host_pipe_destroy: not a pipe");
    hp->magic = 0;
    delete hp;
    return 0;
}
int pomdp_step_packed_host(void* pipe, int kind, const void* params, const void* table, const int32_t* h_state,
                           const int32_t* h_action, int32_t* h_next, int32_t* h_result, int64_t n, int64_t goff, uint64_t seed,
                           uint32_t step) {
    HostPipe* hp = (HostPipe*)pipe;
    int rc = host::check_host_step(hp && hp->magic == 0x50495045u, hp ? hp->words : 0, kind, params, h_state, h_action, h_next,
                                   h_result, n, goff);
    const int W = hp ? hp->words : 1;
    for (int64_t lo = 0; lo < n && rc == 0; lo += hp->chunk) {
        const int64_t m = n - lo < hp->chunk ? n - lo : hp->chunk;
        const int32_t *s = h_state + lo * W, *a = h_action + lo;
        int32_t *nx = h_next + lo * W, *r = h_result + lo;
        switch (kind) {
            case POMDP_KIND_ROCK: rc = pomdp_rock_step_packed((const PomdpRockParams*)params, table, s, a, nx, r, m, goff + lo, seed, step, nullptr); break;
            case POMDP_KIND_TAG: rc = pomdp_tag_step_packed((const PomdpTagParams*)params, table, s, a, nx, r, m, goff + lo, seed, step, nullptr); break;
            case POMDP_KIND_TIGER: rc = pomdp_tiger_step_packed((const PomdpTigerParams*)params, s, a, nx, r, m, goff + lo, seed, step, nullptr); break;
            default: rc = pomdp_network_step_packed((const PomdpNetworkParams*)params, s, a, nx, r, m, goff + lo, seed, step, nullptr); break;
        }
    }
    return rc;
}

int pomdp_rock_belief_update(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* action,
                             const int32_t* obs, int32_t* count, int32_t* measured, double* lkv, double* lkw, double* pv,
                             int64_t n, void*) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if ((rc = host::check_belief(state, action, obs, count, measured, lkv, lkw, pv, n))) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "rock: table is NULL");
    const RockTableHdr* hdr = (const RockTableHdr*)table;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t a = action[i];
        if (a <= 4 || a >= (int32_t)d.n_actions) continue;
        const int64_t j = i * d.k + (a - 5);
        if (host::rock_words(q) == 1) rock_belief_update<uint32_t>(d, hdr, load_state<uint32_t>(state, i), a, obs[i], count[j], measured[j], lkv[j], lkw[j], pv[j]);
        else rock_belief_update<uint64_t>(d, hdr, load_state<uint64_t>(state, i), a, obs[i], count[j], measured[j], lkv[j], lkw[j], pv[j]);
    }
    return 0;
}

int pomdp_rock_legal_list(const PomdpRockParams* q, const void* table, const int32_t* state, uint32_t* list, int64_t n, void*) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if ((rc = host::check_policy(state, list, n, 0, "pomdp_rock_legal_list"))) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "rock: table is NULL");
    const RockLut* lut = (const RockLut*)((const char*)table + ROCK_LUT_OFFSET);
    for (int64_t i = 0; i < n; ++i)
        list[i] = host::rock_words(q) == 1 ? rock_legal_list<uint32_t>(d, lut, load_state<uint32_t>(state, i))
                                           : rock_legal_list<uint64_t>(d, lut, load_state<uint64_t>(state, i));
    return 0;
}
// ---- heuristic action sets and rollouts: the functors the kernels inline (pomdp_core.h / pomdp_envs.h)
int pomdp_rock_history_update(const PomdpRockParams* q, const int32_t* obs_field, const int32_t* action,
                              const int32_t* next_obs_field, int32_t* check_totals, int64_t n, void*) {
    int rc = host::make_rock(q, nullptr, nullptr);
    if (rc) return rc;
    if (n < 0) return host::fail(POMDP_E_BADARG, "pomdp_rock_history_update: n is negative");
    if (n > 0 && (!obs_field || !action || !next_obs_field || !check_totals))
        return host::fail(POMDP_E_BADARG, "pomdp_rock_history_update: a required array pointer is NULL");
    for (int64_t i = 0; i < n; ++i) {
        const int32_t a = action[i];
        if (a < 5 || a >= 5 + q->num_rocks) continue;
        const int64_t j = i * q->num_rocks + (a - 5);
        int32_t ts = rock_totals_sample(check_totals[j]), td = rock_totals_dir(check_totals[j]);
        rock_history_update(a, obs_field[i], next_obs_field[i], ts, td);
        check_totals[j] = rock_totals_pack(ts, td);
    }
    return 0;
}
extern "C++" {
namespace {
template <typename S>
void rock_preferred_host(const RockDev& d, const void* table, const int32_t* state, const int32_t* count, const int32_t* measured,
                         const double* pv, const int32_t* totals, int32_t* out, int64_t n, bool policy, int64_t goff, uint64_t seed,
                         uint32_t step) {
    const RockTableHdr* hdr = (const RockTableHdr*)table;
    const RockLut* lut = (const RockLut*)((const char*)table + ROCK_LUT_OFFSET);
    const PhiloxKey key = philox_key(seed);
    for (int64_t i = 0; i < n; ++i) {
        const RockPlanesView h = {count, measured, pv, totals, i * d.k};
        const S s = load_state<S>(state, i);
        out[i] = policy ? rock_policy_preferred<S>(d, hdr, lut, s, h, draw_word(key, (uint64_t)(goff + i), step, DOMAIN_POLICY, 0))
                        : (int32_t)rock_preferred_mask<S>(d, hdr, s, h);
    }
}
int rock_preferred_entry(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* count,
                         const int32_t* measured, const double* pv, const int32_t* totals, int32_t* out, int64_t n, bool policy,
                         int64_t goff, uint64_t seed, uint32_t step, const char* what) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if ((rc = host::check_policy(state, out, n, goff, what))) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "rock: table is NULL");
    if (host::rock_words(q) == 1) rock_preferred_host<uint32_t>(d, table, state, count, measured, pv, totals, out, n, policy, goff, seed, step);
    else rock_preferred_host<uint64_t>(d, table, state, count, measured, pv, totals, out, n, policy, goff, seed, step);
    return 0;
}
template <typename S, bool STOCH>
void rock_rollout_preferred_host(const RockDev& d, const void* table, const int32_t* state, const int32_t* first_action,
                                 const RockPlanesPtr& pl, int32_t* final_state, double* ret, int32_t* steps, int32_t* flags, int64_t n,
                                 int64_t goff, uint64_t seed, uint32_t step, int32_t max_steps, double gamma, bool next_is_reward) {
    const PhiloxKey key = philox_key(seed);
    const bool records = pl.scratch && !pl.count && !pl.measured && !pl.lkv && !pl.lkw && !pl.pv && !pl.totals && max_steps <= 32767;
    for (int64_t i = 0; i < n; ++i) {
        int32_t prev_ob = pl.prev_obs ? pl.prev_obs[i] : 0;
        S s = load_state<S>(state, i);
        RolloutAcc acc;
        if (records) {
            RockHeurRecords h;
            h.init((RockRec*)pl.scratch + i * d.k);
            rock_rollout_preferred1<S, STOCH>(d, (const unsigned char*)table, s, key, (uint64_t)(goff + i), step, max_steps, gamma,
                                              next_is_reward, first_action != nullptr, first_action ? first_action[i] : 0, h, prev_ob, acc);
        } else {
            RockHeurLocal h;
            h.load(pl, i * d.k, d.k);
            rock_rollout_preferred1<S, STOCH>(d, (const unsigned char*)table, s, key, (uint64_t)(goff + i), step, max_steps, gamma,
                                              next_is_reward, first_action != nullptr, first_action ? first_action[i] : 0, h, prev_ob, acc);
            h.store(pl, i * d.k, d.k);
        }
        if (final_state) store_state(final_state, i, s);
        ret[i] = acc.ret; steps[i] = acc.steps; flags[i] = acc.flags;
        if (pl.prev_obs) pl.prev_obs[i] = prev_ob;
    }
}
}  // namespace
}  // extern "C++"
int pomdp_rock_preferred_mask(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* count,
                              const int32_t* measured, const double* pv, const int32_t* totals, uint32_t* mask, int64_t n, void*) {
    return rock_preferred_entry(q, table, state, count, measured, pv, totals, (int32_t*)mask, n, false, 0, 0, 0, "pomdp_rock_preferred_mask");
}
int pomdp_rock_policy_preferred(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* count,
                                const int32_t* measured, const double* pv, const int32_t* totals, int32_t* action, int64_t n,
                                int64_t goff, uint64_t seed, uint32_t step, void*) {
    return rock_preferred_entry(q, table, state, count, measured, pv, totals, action, n, true, goff, seed, step, "pomdp_rock_policy_preferred");
}
int pomdp_rock_rollout_preferred(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* first_action,
                                 const PomdpRockHeuristicPlanes* planes, int32_t* final_state, double* ret, int32_t* steps,
                                 int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step, int32_t max_steps,
                                 double discount, int32_t next_is_reward, void*) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    if ((rc = host::check_rollout(state, final_state, ret, steps, flags, n, goff, max_steps, "pomdp_rock_rollout_preferred"))) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "rock: table is NULL");
    RockPlanesPtr pl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (planes) { pl.count = planes->count; pl.measured = planes->measured; pl.lkv = planes->lkv; pl.lkw = planes->lkw;
                  pl.pv = planes->prob_valuable; pl.totals = planes->check_totals; pl.prev_obs = planes->prev_obs; pl.scratch = planes->scratch; }
#define HS_RRP(S, ST) rock_rollout_preferred_host<S, ST>(d, table, state, first_action, pl, final_state, ret, steps, flags, n, goff, seed, \
                                                         step, max_steps, discount, next_is_reward != 0)
    if (host::rock_words(q) == 1) { if (d.stochastic) HS_RRP(uint32_t, true); else HS_RRP(uint32_t, false); }
    else { if (d.stochastic) HS_RRP(uint64_t, true); else HS_RRP(uint64_t, false); }
#undef HS_RRP
    return 0;
}
int pomdp_tag_preferred_mask(const PomdpTagParams* q, const void* table, const int32_t* state, const int32_t* last_obs,
                             const int32_t* last_action, uint32_t* mask, int64_t n, void*) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    if ((rc = host::check_policy(state, mask, n, 0, "pomdp_tag_preferred_mask"))) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "tag: table is NULL");
    for (int64_t i = 0; i < n; ++i)
        mask[i] = tag_preferred_mask((const TagTables*)table, (uint32_t)state[i], last_obs ? last_obs[i] : 0, last_action ? last_action[i] : -1);
    return 0;
}
int pomdp_tag_policy_preferred(const PomdpTagParams* q, const void* table, const int32_t* state, const int32_t* last_obs,
                               const int32_t* last_action, int32_t* action, int64_t n, int64_t goff, uint64_t seed, uint32_t step, void*) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    if ((rc = host::check_policy(state, action, n, goff, "pomdp_tag_policy_preferred"))) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "tag: table is NULL");
    const PhiloxKey key = philox_key(seed);
    for (int64_t i = 0; i < n; ++i)
        action[i] = tag_policy_preferred((const TagTables*)table, (uint32_t)state[i], last_obs ? last_obs[i] : 0,
                                         last_action ? last_action[i] : -1, draw_word(key, (uint64_t)(goff + i), step, DOMAIN_POLICY, 0));
    return 0;
}
int pomdp_tag_rollout_preferred(const PomdpTagParams* q, const void* table, const int32_t* state, int32_t* last_obs,
                                int32_t* last_action, const int32_t* first_action, int32_t* final_state, double* ret, int32_t* steps,
                                int32_t* flags, int64_t n, int64_t goff, uint64_t seed, uint32_t step, int32_t max_steps,
                                double discount, void*) {
    TagDev d;
    int rc = host::make_tag(q, &d);
    if (rc) return rc;
    if ((rc = host::check_rollout(state, final_state, ret, steps, flags, n, goff, max_steps, "pomdp_tag_rollout_preferred"))) return rc;
    if (n > 0 && !table) return host::fail(POMDP_E_BADARG, "tag: table is NULL");
    const PhiloxKey key = philox_key(seed);
    for (int64_t i = 0; i < n; ++i) {
        uint32_t s = (uint32_t)state[i];
        int32_t lo = last_obs ? last_obs[i] : 0, la = last_action ? last_action[i] : -1;
        RolloutAcc acc;
        if (d.n_opp == 1) tag_rollout_preferred1<1>(d, (const unsigned char*)table, s, key, (uint64_t)(goff + i), step, max_steps, discount,
                                                     first_action != nullptr, first_action ? first_action[i] : 0, lo, la, acc);
        else tag_rollout_preferred1<4>(d, (const unsigned char*)table, s, key, (uint64_t)(goff + i), step, max_steps, discount,
                                       first_action != nullptr, first_action ? first_action[i] : 0, lo, la, acc);
        if (final_state) final_state[i] = (int32_t)s;
        ret[i] = acc.ret; steps[i] = acc.steps; flags[i] = acc.flags;
        if (last_obs) last_obs[i] = lo;
        if (last_action) last_action[i] = la;
    }
    return 0;
}

// diagnostic of the CUDA library (a memory-traffic probe): nothing to simulate on the host
int pomdp_stream_probe(const int32_t* state, const int32_t* action, int32_t* next, int32_t* obs, float* rw, int32_t* fl, int64_t n,
                       void*) {
    int rc = host::check_io(state, action, next, obs, rw, fl, n);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) { next[i] = state[i] ^ action[i]; obs[i] = action[i]; memcpy(rw + i, state + i, 4); fl[i] = next[i]; }
    return 0;
}
int pomdp_stream_probe_words(int32_t state_words, const int32_t* state, const int32_t* action, int32_t* next, int32_t* obs, float* rw,
                             int32_t* fl, int64_t n, void*) {
    if (state_words == 1) return pomdp_stream_probe(state, action, next, obs, rw, fl, n, nullptr);
    if (state_words != 2) return host::fail(POMDP_E_BADARG, "pomdp_stream_probe_words: state_words must be 1 or 2");
    int rc = host::check_io(state, action, next, obs, rw, fl, n);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        next[2 * i] = state[2 * i] ^ action[i]; next[2 * i + 1] = state[2 * i + 1];
        obs[i] = action[i]; memcpy(rw + i, state + 2 * i, 4); fl[i] = state[2 * i + 1];
    }
    return 0;
}
int pomdp_coord_op(int32_t op, int32_t xs, int32_t ys, const int32_t* a, const int32_t* b, int32_t* out, int64_t n, void*) {
    const int rc = host::check_coord_op(op, xs, a, b, out, n);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        switch (op) {
            case POMDP_COORD_GET_INDEX: out[i] = grid_get_index(xs, a[2 * i], a[2 * i + 1]); break;
            case POMDP_COORD_GET_COORD: out[2 * i] = a[i] % xs; out[2 * i + 1] = a[i] / xs; break;
            case POMDP_COORD_IS_INSIDE: out[i] = grid_is_inside(xs, ys, a[2 * i], a[2 * i + 1]); break;
            case POMDP_COORD_ADD_MOVE: out[2 * i] = a[2 * i] + move_dx(b[i]); out[2 * i + 1] = a[2 * i + 1] + move_dy(b[i]); break;
            case POMDP_COORD_L1: out[i] = l1_distance(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]); break;
            case POMDP_COORD_TAG_GET_INDEX: out[i] = tag_is_inside(a[2 * i], a[2 * i + 1]) ? tag_get_index(a[2 * i], a[2 * i + 1]) : -1; break;
            case POMDP_COORD_TAG_GET_COORD: {
                int x = -1, y = -1;
                if ((uint32_t)a[i] < (uint32_t)TAG_CELLS) tag_get_coord((uint32_t)a[i], x, y);
                out[2 * i] = x; out[2 * i + 1] = y;
                break;
            }
            case POMDP_COORD_TAG_IS_INSIDE: out[i] = tag_is_inside(a[2 * i], a[2 * i + 1]); break;
            default: return host::fail(POMDP_E_BADARG, "coord: unknown op %d", op);
        }
    }
    return 0;
}

int pomdp_belief_hist_bins(int32_t kind, int32_t p0, int32_t p1) { return host::hist_bins(kind, p0, p1); }
int pomdp_belief_hist(int32_t kind, int32_t p0, int32_t p1, const int32_t* state, int32_t words, int64_t n,
                      long long* hist, void*) {
    const int rc = host::check_hist(kind, p0, p1, state, words, n, hist, 512);
    if (rc) return rc;
    for (int64_t i = 0; i < n; ++i) {
        uint32_t s[SHIP_WORDS] = {0};
        for (int k = 0; k < words && k < SHIP_WORDS; ++k) s[k] = (uint32_t)state[i * words + k];
        belief_bins(kind, p0, p1, s, [&](int bin) { ++hist[bin]; });
    }
    return 0;
}

int pomdp_belief_hist_once(int32_t kind, int32_t p0, int32_t p1, const int32_t* state, int32_t words, int64_t n,
                           long long* scratch, long long* hist_out, void*) {
    int rc = host::check_hist(kind, p0, p1, state, words, n, scratch, POMDP_HIST_MAX_BINS);
    if (rc) return rc;
    if (!scratch) return host::fail(POMDP_E_BADARG, "pomdp_belief_hist_once: scratch is NULL");
    if (!hist_out || ((uintptr_t)hist_out & 7)) return host::fail(POMDP_E_BADARG, "pomdp_belief_hist_once: hist_out is NULL or not 8-byte aligned");
    const int bins = host::hist_bins(kind, p0, p1);
    rc = pomdp_belief_hist(kind, p0, p1, state, words, n, scratch, nullptr);
    if (rc) return rc;
    for (int b = 0; b < bins; ++b) { hist_out[b] = scratch[b]; scratch[b] = 0; }
    return 0;
}

// host stand-in of the fused histogram + all-reduce: the peer table holds host pointers; arrivals are counted, nobody waits
// (the "ranks" of a test call one after the other), hist_out receives the slot as it stands after this rank's additions
int pomdp_belief_hist_allreduce(int32_t kind, int32_t p0, int32_t p1, const int32_t* state, int32_t words, int64_t n,
                                long long* scratch, const void* const* peer_bufs, int32_t world, int32_t rank, int32_t wait,
                                long long* hist_out, void*) {
    int rc = host::check_hist(kind, p0, p1, state, words, n, scratch, POMDP_HIST_MAX_BINS);
    if (rc) return rc;
    if (!scratch || !peer_bufs || world < 1 || world > POMDP_HIST_MAX_RANKS || rank < 0 || rank >= world) return host::fail(POMDP_E_BADARG, "bad peer table");
    const int bins = host::hist_bins(kind, p0, p1);
    rc = pomdp_belief_hist(kind, p0, p1, state, words, n, scratch, nullptr);
    if (rc) return rc;
    const long long epoch = scratch[bins + 1] + 1;
    const int slot = (int)((epoch - 1) & 1);
    long long* own = (long long*)peer_bufs[rank];
    memset(own + (slot ^ 1) * POMDP_HIST_MAX_BINS, 0, POMDP_HIST_MAX_BINS * sizeof(long long));
    for (int b = 0; b < bins; ++b) {
        for (int r = 0; r < world; ++r) ((long long*)peer_bufs[r])[slot * POMDP_HIST_MAX_BINS + b] += scratch[b];
        scratch[b] = 0;
    }
    if (wait)
        for (int r = 0; r < world; ++r) ((long long*)peer_bufs[r])[2 * POMDP_HIST_MAX_BINS + rank] += 1;
    if (hist_out) memcpy(hist_out, own + slot * POMDP_HIST_MAX_BINS, bins * sizeof(long long));
    scratch[bins + 1] = epoch;
    return 0;
}

// host stand-ins of the step kernels with the histogram epilogue: the step, then the counts of next_state through the sink
static int hist_sink_host(int kind, int p0, int p1, const int32_t* next, int words, int64_t n, const PomdpHistSink* sink, const char* what) {
    if (!sink || !sink->scratch || ((uintptr_t)sink->scratch & 7) || ((uintptr_t)sink->hist_out & 7))
        return host::fail(POMDP_E_BADARG, "%s: the histogram sink needs an 8-byte aligned scratch (and hist_out)", what);
    if (sink->d_peer_bufs)
        return pomdp_belief_hist_allreduce(kind, p0, p1, next, words, n, sink->scratch, sink->d_peer_bufs, sink->world, sink->rank,
                                           sink->wait, sink->hist_out, nullptr);
    if (!sink->hist_out) return host::fail(POMDP_E_BADARG, "%s: a local histogram sink (no peer table) needs hist_out", what);
    return pomdp_belief_hist_once(kind, p0, p1, next, words, n, sink->scratch, sink->hist_out, nullptr);
}
int pomdp_rock_step_hist(const PomdpRockParams* q, const void* table, const int32_t* state, const int32_t* action, int32_t* next,
                         int32_t* obs, float* rw, int32_t* fl, int64_t n, int64_t goff, uint64_t seed, uint32_t step,
                         const PomdpHistSink* sink, void*) {
    int rc = pomdp_rock_step(q, table, state, action, next, obs, rw, fl, n, goff, seed, step, nullptr);
    if (rc) return rc;
    return hist_sink_host(POMDP_KIND_ROCK, q->num_rocks, host::rock_words(q), next, host::rock_words(q), n, sink, "pomdp_rock_step_hist");
}
int pomdp_tag_step_hist(const PomdpTagParams* q, const void* table, const int32_t* state, const int32_t* action, int32_t* next,
                        int32_t* obs, float* rw, int32_t* fl, int64_t n, int64_t goff, uint64_t seed, uint32_t step,
                        const PomdpHistSink* sink, void*) {
    int rc = pomdp_tag_step(q, table, state, action, next, obs, rw, fl, n, goff, seed, step, nullptr);
    if (rc) return rc;
    return hist_sink_host(POMDP_KIND_TAG, 0, 0, next, 1, n, sink, "pomdp_tag_step_hist");
}
int pomdp_tiger_step_hist(const PomdpTigerParams* q, const int32_t* state, const int32_t* action, int32_t* next, int32_t* obs,
                          float* rw, int32_t* fl, int64_t n, int64_t goff, uint64_t seed, uint32_t step, const PomdpHistSink* sink,
                          void*) {
    int rc = pomdp_tiger_step(q, state, action, next, obs, rw, fl, n, goff, seed, step, nullptr);
    if (rc) return rc;
    return hist_sink_host(POMDP_KIND_TIGER, 0, 0, next, 1, n, sink, "pomdp_tiger_step_hist");
}
int pomdp_network_step_hist(const PomdpNetworkParams* q, const int32_t* state, const int32_t* action, int32_t* next, int32_t* obs,
                            float* rw, int32_t* fl, int64_t n, int64_t goff, uint64_t seed, uint32_t step, const PomdpHistSink* sink,
                            void*) {
    int rc = pomdp_network_step(q, state, action, next, obs, rw, fl, n, goff, seed, step, nullptr);
    if (rc) return rc;
    return hist_sink_host(POMDP_KIND_NETWORK, q->n_machines, 0, next, 1, n, sink, "pomdp_network_step_hist");
}

// test-only: the division-free float32 reward conversion of network_step_n, for the exhaustive check in
// tests/test_edge_cases.py
void pomdp_hostsim_tenths_to_float(const int32_t* t, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = tenths_to_float(t[i]);
}

// test-only: the one-LOP3 form of the sixteen reset status codes, for the brute-force check against the per-rock
// definition (rock_status_code(rock_reset_word(w, i))) in tests/test_edge_cases.py
void pomdp_hostsim_rock_reset_codes16(const uint32_t* w, uint32_t* fast, uint32_t* per_rock, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        fast[i] = rock_reset_codes16(w[i]);
        uint32_t c = 0;
        for (int r = 0; r < 16; ++r) c |= rock_status_code(rock_reset_word(w[i], r)) << (2 * r);
        per_rock[i] = c;
    }
}

// test-only: the four-env reset path of the vector kernel (shift + LOP3 per env, ties out of line) on explicit draw
// words, next to the per-env definition -- the kernels can only be fed Philox words, which never tie
int pomdp_hostsim_rock_reset4(const PomdpRockParams* q, const uint32_t* words, uint64_t* fast, uint64_t* slow, int64_t n_groups) {
    RockDev d;
    int rc = host::make_rock(q, &d, nullptr);
    if (rc) return rc;
    for (int64_t g = 0; g < n_groups; ++g) {
        const U4 w = {words[4 * g], words[4 * g + 1], words[4 * g + 2], words[4 * g + 3]};
        if (host::rock_words(q) == 1) {
            uint32_t o[4];
            rock_reset4_words<uint32_t>(d, w, o);
            for (int j = 0; j < 4; ++j) { fast[4 * g + j] = o[j]; slow[4 * g + j] = rock_reset_from_word<uint32_t>(d, words[4 * g + j]); }
        } else {
            uint64_t o[4];
            rock_reset4_words<uint64_t>(d, w, o);
            for (int j = 0; j < 4; ++j) { fast[4 * g + j] = o[j]; slow[4 * g + j] = rock_reset_from_word<uint64_t>(d, words[4 * g + j]); }
        }
    }
    return 0;
}

}  // extern "C"
