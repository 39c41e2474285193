// pomdp_envs.h -- per-env adapters between the transition functors of pomdp_core.h and the generic streaming
// kernels of pomdp_kernels.cu (step / reset / policy / rollout).  `__host__ __device__` like pomdp_core.h, so
// tests/hostsim/ drives exactly the same adapters on the CPU.
#pragma once
#include "pomdp_core.h"

namespace pomdp {

// ------------------------------------------------------------------- env policies ---
// Each policy adapts one env's functors from pomdp_core.h to the generic streaming kernels:
//   step4 / reset4 : the FOUR envs of one aligned draw group (one thread, one Philox call per slot)
//   step1 / reset1 : a single env (tails, unaligned views, masked resets)
//   policy         : np.random.choice(env._generate_legal()) from one draw word (uniform over the reference's list)
//   is_done        : the packed state's terminal bit
template <typename S, bool STOCH>
struct RockEnvT {
    typedef RockDev Params;
    typedef S State;
    static constexpr int kHistKind = 0;              // POMDP_KIND_ROCK: the belief-histogram bins of this env's states
    static POMDP_HD int hist_p0(const Params& p) { return p.k; }
    static constexpr bool kTable = true;             // table built on the host, staged per CTA by one TMA bulk copy
    static constexpr bool kParamTable = false;
    static POMDP_HD int32_t policy(const Params& p, const unsigned char* tbl, S s, uint32_t w) {   // rock.py:273-291
        return rock_policy<S>(p, reinterpret_cast<const RockTableHdr*>(tbl),
                              reinterpret_cast<const RockLut*>(tbl + ROCK_LUT_OFFSET), s, w);
    }
    static POMDP_HD bool is_done(S s) { return (s & RockBits<S>::DONE) != 0; }
    static POMDP_HD double obs_prob(const Params& p, const unsigned char* tbl, S s, int32_t a, int32_t ob, double) {
        return rock_obs_prob<S>(p, reinterpret_cast<const RockTableHdr*>(tbl), s, a, ob);
    }
    static POMDP_HD void legal_mask(const Params& p, const unsigned char* tbl, S s, uint32_t* m) {
        m[0] = rock_legal_mask<S>(p, reinterpret_cast<const RockTableHdr*>(tbl), reinterpret_cast<const RockLut*>(tbl + ROCK_LUT_OFFSET), s);
    }
    static POMDP_HD int mask_words(const Params&) { return 1; }
    static POMDP_HD double reward64(float rw) { return (double)rw; }          // integral rewards (rock.py:141-169)
    static POMDP_HD int32_t reward_units(float rw) { return reward_units_int(rw); }
    static POMDP_HD void step4(const Params& p, const unsigned char* tbl, const S s[4], const int32_t a[4],
                                                 const PhiloxKey& seed, uint64_t group, uint32_t ctr, S s2[4],
                                                 int32_t ob[4], float rw[4], int32_t fl[4]) {
        const RockRes* rtab = reinterpret_cast<const RockRes*>(tbl + ROCK_RTAB_OFFSET);
        const RockLut* lut = reinterpret_cast<const RockLut*>(tbl + ROCK_LUT_OFFSET);
        const U4 qs = draw_quad(seed, group, ctr, DOMAIN_STEP, 1);
        U4 qg = {0, 0, 0, 0};
        if (STOCH) qg = draw_quad(seed, group, ctr, DOMAIN_STEP, 0);
        rock_step<S, STOCH>(p, lut, rtab, s[0], a[0], qg.x, qs.x, s2[0], ob[0], rw[0], fl[0]);
        rock_step<S, STOCH>(p, lut, rtab, s[1], a[1], qg.y, qs.y, s2[1], ob[1], rw[1], fl[1]);
        rock_step<S, STOCH>(p, lut, rtab, s[2], a[2], qg.z, qs.z, s2[2], ob[2], rw[2], fl[2]);
        rock_step<S, STOCH>(p, lut, rtab, s[3], a[3], qg.w, qs.w, s2[3], ob[3], rw[3], fl[3]);
    }
    static POMDP_HD void step1(const Params& p, const unsigned char* tbl, S s, int32_t a,
                                                 const PhiloxKey& seed, uint64_t env, uint32_t ctr, S& s2, int32_t& ob,
                                                 float& rw, int32_t& fl) {
        const RockRes* rtab = reinterpret_cast<const RockRes*>(tbl + ROCK_RTAB_OFFSET);
        const RockLut* lut = reinterpret_cast<const RockLut*>(tbl + ROCK_LUT_OFFSET);
        const uint32_t ws = draw_word(seed, env, ctr, DOMAIN_STEP, 1);
        const uint32_t wg = STOCH ? draw_word(seed, env, ctr, DOMAIN_STEP, 0) : 0u;
        rock_step<S, STOCH>(p, lut, rtab, s, a, wg, ws, s2, ob, rw, fl);
    }
    static POMDP_HD void reset4(const Params& p, const PhiloxKey& seed, uint64_t group, uint32_t ctr,
                                                  S s[4], int32_t ob[4]) {
        rock_reset4<S>(p, seed, group, ctr, s);
        ob[0] = ob[1] = ob[2] = ob[3] = 0;
    }
    static POMDP_HD void reset1(const Params& p, const PhiloxKey& seed, uint64_t env, uint32_t ctr,
                                                  S& s, int32_t& ob) {
        s = rock_reset<S>(p, LazyDraw{&seed, env, ctr, DOMAIN_RESET});
        ob = 0;
    }
};

// Precomputes the NS draw words of each of the four envs of a group (NS Philox calls).
template <int NS>
POMDP_HD void quad_words(const PhiloxKey& seed, uint64_t group, uint32_t ctr, uint32_t domain, int n_used,
                                           WordDraw<NS> d[4]) {
    POMDP_UNROLL
    for (int slot = 0; slot < NS; ++slot) {
        U4 q = {0, 0, 0, 0};
        if (slot < n_used) q = draw_quad(seed, group, ctr, domain, (uint32_t)slot);   // uniform branch
        d[0].w[slot] = q.x; d[1].w[slot] = q.y; d[2].w[slot] = q.z; d[3].w[slot] = q.w;
    }
}

// NOPP = 1: the stock Tag-v0 (one draw slot); NOPP = 4: any num_opponents in 1..4.
template <int NOPP>
struct TagEnvT {
    typedef TagDev Params;
    typedef uint32_t State;
    static constexpr int kHistKind = 1;              // POMDP_KIND_TAG
    static POMDP_HD int hist_p0(const Params&) { return 0; }
    static constexpr bool kTable = true;             // TagTables: built on the host, staged per CTA by one TMA bulk copy
    static constexpr bool kParamTable = false;
    static POMDP_HD int32_t policy(const Params&, const unsigned char*, State, uint32_t w) {       // tag.py:228-229
        return (int32_t)rand_below(w, 5u);
    }
    static POMDP_HD bool is_done(State s) { return (s & TAG_DONE) != 0; }
    static POMDP_HD double obs_prob(const Params& p, const unsigned char*, State s, int32_t, int32_t ob, double) {
        return tag_obs_prob(p, s, ob);
    }
    static POMDP_HD void legal_mask(const Params&, const unsigned char*, State, uint32_t* m) { m[0] = 31u; }   // tag.py:228-229
    static POMDP_HD int mask_words(const Params&) { return 1; }
    static POMDP_HD double reward64(float rw) { return (double)rw; }
    static POMDP_HD int32_t reward_units(float rw) { return reward_units_int(rw); }
    // kFixUp = false (the step kernels): a group with a flagged env -- a done state, an action > 4, a cell id off the
    // board: everything the reference asserts on, ONE test per group -- leaves for the checked functor, so the common
    // path carries no fix-up code.  kFixUp = true (the rollouts, where finished episodes make such groups common): the
    // transition is formed for all four first and flagged envs are patched afterwards.  Same results; measured both ways
    // (profiles/r05b, r05c: step 17.5 vs 18.4 us at 2^22, rollout 636 vs 528 us).
    template <bool kFixUp>
    static POMDP_HD void step4_1opp(const Params& p, const TagTables* T, const State s[4], const int32_t a[4], const U4& q,
                                    State s2[4], int32_t ob[4], float rw[4], int32_t fl[4]) {
        const uint32_t e0 = tag_lut_word(T, s[0], a[0]), e1 = tag_lut_word(T, s[1], a[1]), e2 = tag_lut_word(T, s[2], a[2]),
                       e3 = tag_lut_word(T, s[3], a[3]);
        const uint32_t amax = umax32(umax32((uint32_t)a[0], (uint32_t)a[1]), umax32((uint32_t)a[2], (uint32_t)a[3]));
        const uint32_t emin = umin32(umin32(e0, e1), umin32(e2, e3));
        const bool flagged = (int32_t)(s[0] | s[1] | s[2] | s[3]) < 0 || amax > 4u || emin == 0u;
        if (!kFixUp && flagged) {
            tag_step_1opp(p, T, s[0], a[0], q.x, s2[0], ob[0], rw[0], fl[0]);
            tag_step_1opp(p, T, s[1], a[1], q.y, s2[1], ob[1], rw[1], fl[1]);
            tag_step_1opp(p, T, s[2], a[2], q.z, s2[2], ob[2], rw[2], fl[2]);
            tag_step_1opp(p, T, s[3], a[3], q.w, s2[3], ob[3], rw[3], fl[3]);
            return;
        }
        tag_step_1opp_fast(p, e0, s[0], q.x, s2[0], ob[0], rw[0], fl[0]);
        tag_step_1opp_fast(p, e1, s[1], q.y, s2[1], ob[1], rw[1], fl[1]);
        tag_step_1opp_fast(p, e2, s[2], q.z, s2[2], ob[2], rw[2], fl[2]);
        tag_step_1opp_fast(p, e3, s[3], q.w, s2[3], ob[3], rw[3], fl[3]);
        if (kFixUp && flagged) {
            const uint32_t e[4] = {e0, e1, e2, e3};
            POMDP_UNROLL
            for (int j = 0; j < 4; ++j) {                        // tag.py:109-110, 116-117: flagged, state untouched, obs = reward = 0
                const int32_t err = (s[j] & TAG_DONE) ? (int32_t)(FLAG_DONE | FLAG_STEPPED_DONE)
                                    : ((uint32_t)a[j] >= 5u) ? (int32_t)FLAG_BAD_ACTION
                                    : (e[j] == 0u) ? (int32_t)FLAG_BAD_STATE : 0;
                if (err) { s2[j] = s[j]; ob[j] = 0; rw[j] = 0.f; fl[j] = err; }
            }
        }
    }
    static POMDP_HD void step4(const Params& p, const unsigned char* tbl, const State s[4], const int32_t a[4],
                                                 const PhiloxKey& seed, uint64_t group, uint32_t ctr, State s2[4], int32_t ob[4],
                                                 float rw[4], int32_t fl[4]) {
        const TagTables* T = reinterpret_cast<const TagTables*>(tbl);
        if (NOPP == 1) {
            step4_1opp<false>(p, T, s, a, draw_quad(seed, group, ctr, DOMAIN_STEP, 0), s2, ob, rw, fl);
            return;
        }
        tag_step4_multi(p, T, s, a, seed, group, ctr, s2, ob, rw, fl);
    }
    // the rollout kernels' call: the same step, fix-up flavour
    static POMDP_HD void step4_rollout(const Params& p, const unsigned char* tbl, const State s[4], const int32_t a[4],
                                       const PhiloxKey& seed, uint64_t group, uint32_t ctr, State s2[4], int32_t ob[4],
                                       float rw[4], int32_t fl[4]) {
        const TagTables* T = reinterpret_cast<const TagTables*>(tbl);
        if (NOPP == 1) {
            step4_1opp<true>(p, T, s, a, draw_quad(seed, group, ctr, DOMAIN_STEP, 0), s2, ob, rw, fl);
            return;
        }
        tag_step4_multi(p, T, s, a, seed, group, ctr, s2, ob, rw, fl);
    }
    static POMDP_HD void step1(const Params& p, const unsigned char* tbl, State s, int32_t a, const PhiloxKey& seed,
                                                 uint64_t env, uint32_t ctr, State& s2, int32_t& ob, float& rw,
                                                 int32_t& fl) {
        tag_step(p, reinterpret_cast<const TagTables*>(tbl), s, a, LazyDraw{&seed, env, ctr, DOMAIN_STEP}, s2, ob, rw, fl);
    }
    static POMDP_HD void reset4(const Params& p, const PhiloxKey& seed, uint64_t group, uint32_t ctr,
                                                  State s[4], int32_t ob[4]) {
        constexpr int NS = (1 + NOPP + TAG_DIGITS_PER_WORD - 1) / TAG_DIGITS_PER_WORD;     // three cells per draw word
        WordDraw<NS> d[4];
        quad_words<NS>(seed, group, ctr, DOMAIN_RESET, (1 + p.n_opp + TAG_DIGITS_PER_WORD - 1) / TAG_DIGITS_PER_WORD, d);
        POMDP_UNROLL
        for (int j = 0; j < 4; ++j) tag_reset(p, d[j], s[j], ob[j]);
    }
    static POMDP_HD void reset1(const Params& p, const PhiloxKey& seed, uint64_t env, uint32_t ctr, State& s,
                                                  int32_t& ob) {
        tag_reset(p, LazyDraw{&seed, env, ctr, DOMAIN_RESET}, s, ob);
    }
};

// Tag's table-free queries (observation likelihood, legal mask)
struct TagNoTable {
    typedef TagDev Params;
    typedef uint32_t State;
    static constexpr bool kTable = false;
    static constexpr bool kParamTable = false;
    static POMDP_HD double obs_prob(const Params& p, const unsigned char*, State s, int32_t, int32_t ob, double) {
        return tag_obs_prob(p, s, ob);
    }
    static POMDP_HD void legal_mask(const Params&, const unsigned char*, State, uint32_t* m) { m[0] = 31u; }   // tag.py:228-229
    static POMDP_HD int mask_words(const Params&) { return 1; }
};

struct TigerEnvP {
    typedef TigerDev Params;
    typedef uint32_t State;
    static constexpr int kHistKind = 3;              // POMDP_KIND_TIGER
    static POMDP_HD int hist_p0(const Params&) { return 0; }
    static constexpr bool kTable = false;
    static constexpr bool kParamTable = false;
    static POMDP_HD int32_t policy(const Params&, const unsigned char*, State, uint32_t w) {       // tiger.py:111-112
        return (int32_t)rand_below(w, 3u);
    }
    static POMDP_HD bool is_done(State s) { return (s & TIGER_DONE) != 0; }
    static POMDP_HD double obs_prob(const Params&, const unsigned char*, State s, int32_t a, int32_t ob, double correct_prob) {
        return tiger_obs_prob(correct_prob, s, a, ob);
    }
    static POMDP_HD void legal_mask(const Params&, const unsigned char*, State, uint32_t* m) { m[0] = 7u; }    // tiger.py:111-112
    static POMDP_HD int mask_words(const Params&) { return 1; }
    static POMDP_HD double reward64(float rw) { return (double)rw; }
    static POMDP_HD int32_t reward_units(float rw) { return reward_units_int(rw); }
    static POMDP_HD void step4(const Params& p, const unsigned char*, const State s[4], const int32_t a[4],
                                                 const PhiloxKey& seed, uint64_t group, uint32_t ctr, State s2[4], int32_t ob[4],
                                                 float rw[4], int32_t fl[4]) {
        WordDraw<1> d[4];
        quad_words<1>(seed, group, ctr, DOMAIN_STEP, 1, d);
    POMDP_UNROLL
        for (int j = 0; j < 4; ++j) tiger_step(p, s[j], a[j], d[j], s2[j], ob[j], rw[j], fl[j]);
    }
    static POMDP_HD void step1(const Params& p, const unsigned char*, State s, int32_t a, const PhiloxKey& seed,
                                                 uint64_t env, uint32_t ctr, State& s2, int32_t& ob, float& rw,
                                                 int32_t& fl) {
        tiger_step(p, s, a, LazyDraw{&seed, env, ctr, DOMAIN_STEP}, s2, ob, rw, fl);
    }
    static POMDP_HD void reset4(const Params&, const PhiloxKey& seed, uint64_t group, uint32_t ctr, State s[4],
                                                  int32_t ob[4]) {
        WordDraw<1> d[4];
        quad_words<1>(seed, group, ctr, DOMAIN_RESET, 1, d);
    POMDP_UNROLL
        for (int j = 0; j < 4; ++j) tiger_reset(d[j], s[j], ob[j]);
    }
    static POMDP_HD void reset1(const Params&, const PhiloxKey& seed, uint64_t env, uint32_t ctr, State& s,
                                                  int32_t& ob) {
        tiger_reset(LazyDraw{&seed, env, ctr, DOMAIN_RESET}, s, ob);
    }
};

// G: compile-time number of five-machine groups for the step (0 = runtime, network_step_n)
template <int G>
struct NetworkEnvT {
    typedef NetworkDev Params;
    typedef uint32_t State;
    static constexpr int kHistKind = 4;              // POMDP_KIND_NETWORK
    static POMDP_HD int hist_p0(const Params& p) { return p.n; }
    static constexpr bool kTable = false;
    // the step's tables (alias columns of the joint failure draw, neighbour-down map: 2.8 KB) travel inside the kernel
    // parameters and are copied to shared memory once per CTA by the kernels that step (step, rollout)
    static constexpr bool kParamTable = true;
    static constexpr uint32_t kParamTableBytes = (uint32_t)sizeof(NetworkTables);
    static POMDP_HD const void* param_table(const Params& p) { return &p.t; }
#if defined(__CUDA_ARCH__)
    typedef NetTabSmem Tab;
    static __device__ __forceinline__ Tab tables(const Params&, const unsigned char* tbl) {      // staged in shared memory
        uint32_t base;                                           // opaque to the optimiser, or it is re-derived per lookup
        asm("mov.u32 %0, %1;" : "=r"(base) : "r"((uint32_t)__cvta_generic_to_shared(tbl)));
        return Tab{base};
    }
#else
    typedef NetTabPtr Tab;
    static POMDP_HD Tab tables(const Params& p, const unsigned char* tbl) {
        return Tab{tbl ? reinterpret_cast<const NetworkTables*>(tbl) : &p.t};
    }
#endif
    static POMDP_HD int32_t policy(const Params& p, const unsigned char*, State, uint32_t w) {     // network.py:129-130
        return (int32_t)rand_below(w, (uint32_t)(2 * p.n + 1));
    }
    static POMDP_HD bool is_done(State s) { return (s & NETWORK_DONE) != 0; }
    static POMDP_HD double obs_prob(const Params& p, const unsigned char*, State s, int32_t a, int32_t ob, double) {
        return network_obs_prob(p, p.p_ob, s, a, ob);
    }
    static POMDP_HD void legal_mask(const Params& p, const unsigned char*, State, uint32_t* m) {              // network.py:129-130
        const int na = 2 * p.n + 1;
        m[0] = na >= 32 ? 0xFFFFFFFFu : ((1u << na) - 1u);
        if (na > 32) m[1] = (1u << (na - 32)) - 1u;
    }
    static POMDP_HD int mask_words(const Params& p) { return (2 * p.n + 1 + 31) / 32; }
    // the reference's reward is the Python double s - 0.1 / s - 2.5 / s == tenths / 10.0 (network.py:87-108); the
    // float32 the step kernel emits determines the integer number of tenths uniquely
    static POMDP_HD int32_t reward_units(float rw) { return reward_units_tenths(rw); }
    static POMDP_HD double reward64(float rw) {
        const double t = (double)rw * 10.0;
        return (double)(long long)(t < 0 ? t - 0.5 : t + 0.5) / 10.0;
    }
    static POMDP_HD void step4(const Params& p, const unsigned char* tbl, const State s[4], const int32_t a[4],
                                                 const PhiloxKey& seed, uint64_t group, uint32_t ctr, State s2[4], int32_t ob[4],
                                                 float rw[4], int32_t fl[4]) {
        network_step_n<4, Tab, G>(p, tables(p, tbl), s, a, seed, group, 0, ctr, s2, ob, rw, fl);
    }
    static POMDP_HD void step1(const Params& p, const unsigned char* tbl, State s, int32_t a, const PhiloxKey& seed,
                                                 uint64_t env, uint32_t ctr, State& s2, int32_t& ob, float& rw,
                                                 int32_t& fl) {
        network_step_n<1, Tab, G>(p, tables(p, tbl), &s, &a, seed, env >> 2, (int)(env & 3), ctr, &s2, &ob, &rw, &fl);
    }
    static POMDP_HD void reset4(const Params& p, const PhiloxKey&, uint64_t, uint32_t, State s[4],
                                                  int32_t ob[4]) {
        s[0] = s[1] = s[2] = s[3] = (1u << p.n) - 1u;   // network.py:61-69: all up, obs = OFF (0)
        ob[0] = ob[1] = ob[2] = ob[3] = 0;
    }
    static POMDP_HD void reset1(const Params& p, const PhiloxKey&, uint64_t, uint32_t, State& s, int32_t& ob) {
        s = (1u << p.n) - 1u;
        ob = 0;
    }
};
typedef NetworkEnvT<0> NetworkEnvP;      // any n_machines
typedef NetworkEnvT<2> NetworkEnv10;     // 6..10 machines (the stock Network-v0): two groups


// One env through a whole rollout (scalar kernel path; tests/hostsim runs the same function).
template <class Env>
POMDP_HD void rollout1(const typename Env::Params& p, const unsigned char* tbl, typename Env::State& s,
                       const PhiloxKey& seed, uint64_t env, uint32_t ctr0, int32_t max_steps, double gamma,
                       RolloutAcc& acc, bool has_first = false, int32_t first_action = 0) {
    acc.init(Env::is_done(s));
    for (int32_t t = 0; t < max_steps && !Env::is_done(s); ++t) {
        const uint32_t ctr = ctr0 + (uint32_t)t;
        const int32_t a = (t == 0 && has_first) ? first_action
                                                : Env::policy(p, tbl, s, draw_word(seed, env, ctr, DOMAIN_POLICY, 0));
        typename Env::State s2;
        int32_t ob, fl;
        float rw;
        Env::step1(p, tbl, s, a, seed, env, ctr, s2, ob, rw, fl);
        s = s2;
        acc.add(Env::reward64(rw), gamma, fl);
    }
}

// ---- heuristic policies and rollouts: shared by the kernels and tests/hostsim ----------------
// Read-only view of one env's rows of the per-rock planes; a NULL plane stands for its fresh value.
struct RockPlanesView {
    const int32_t* cnt_p; const int32_t* meas_p; const double* pv_p; const int32_t* tot_p;
    int64_t base;
    POMDP_HD int32_t tot_sample(int i) const { return tot_p ? rock_totals_sample(tot_p[base + i]) : 0; }
    POMDP_HD int32_t tot_dir(int i) const { return tot_p ? rock_totals_dir(tot_p[base + i]) : 0; }
    POMDP_HD int32_t count(int i) const { return cnt_p ? cnt_p[base + i] : 0; }
    POMDP_HD int32_t measured(int i) const { return meas_p ? meas_p[base + i] : 0; }
    POMDP_HD double pv(int i) const { return pv_p ? pv_p[base + i] : .5; }
    // the rule's three per-rock predicates as bit masks (rock_preferred_mask)
    POMDP_HD uint32_t sample_mask(int k) const {
        uint32_t m = 0;
        for (int i = 0; i < k; ++i) m |= (rock_pred_sample(tot_sample(i)) ? 1u : 0u) << i;
        return m;
    }
    POMDP_HD uint32_t dir_mask(int k) const {
        uint32_t m = 0;
        for (int i = 0; i < k; ++i) m |= (rock_pred_dir(tot_dir(i)) ? 1u : 0u) << i;
        return m;
    }
    POMDP_HD uint32_t check_mask(int k) const {
        uint32_t m = 0;
        for (int i = 0; i < k; ++i) m |= (rock_pred_check(measured(i), count(i), pv(i)) ? 1u : 0u) << i;
        return m;
    }
};
struct RockPlanesPtr { int32_t* count; int32_t* measured; double* lkv; double* lkw; double* pv; int32_t* totals; int32_t* prev_obs; void* scratch; };
// One env's planes for the length of a rollout (local memory on the device: k <= 16 rocks x 40 B)
struct RockHeurLocal {
    int32_t cnt[16], meas[16], ts[16], td[16];
    double lkv[16], lkw[16], pvv[16];
    uint32_t m_sample, m_dir, m_check;      // the rule's per-rock predicates, kept current: registers, not local memory
    POMDP_HD uint32_t sample_mask(int) const { return m_sample; }
    POMDP_HD uint32_t dir_mask(int) const { return m_dir; }
    POMDP_HD uint32_t check_mask(int) const { return m_check; }
    POMDP_HD void refresh(int r) {          // after rock r's planes changed (it was checked)
        const uint32_t bit = 1u << r;
        m_sample = (m_sample & ~bit) | (rock_pred_sample(ts[r]) ? bit : 0u);
        m_dir = (m_dir & ~bit) | (rock_pred_dir(td[r]) ? bit : 0u);
        m_check = (m_check & ~bit) | (rock_pred_check(meas[r], cnt[r], pvv[r]) ? bit : 0u);
    }
    template <typename S>
    POMDP_HD void check(const RockDev& p, const RockTableHdr* hdr, S s, int32_t a, int32_t ob, int32_t obs_field, int32_t next_field) {
        const int r = a - 5;
        rock_belief_update<S>(p, hdr, s, a, ob, cnt[r], meas[r], lkv[r], lkw[r], pvv[r]);   // rock.py:177-191
        rock_history_update(a, obs_field, next_field, ts[r], td[r]);                        // rock.py:566
        refresh(r);
    }
    POMDP_HD void load(const RockPlanesPtr& pl, int64_t base, int k) {
        m_sample = m_dir = m_check = 0u;
        for (int r = 0; r < k; ++r) {
            cnt[r] = pl.count ? pl.count[base + r] : 0;
            meas[r] = pl.measured ? pl.measured[base + r] : 0;
            lkv[r] = pl.lkv ? pl.lkv[base + r] : 1.0;
            lkw[r] = pl.lkw ? pl.lkw[base + r] : 1.0;
            pvv[r] = pl.pv ? pl.pv[base + r] : .5;
            const int32_t t = pl.totals ? pl.totals[base + r] : 0;
            ts[r] = rock_totals_sample(t);
            td[r] = rock_totals_dir(t);
            refresh(r);
        }
    }
    POMDP_HD void store(const RockPlanesPtr& pl, int64_t base, int k) const {
        for (int r = 0; r < k; ++r) {
            if (pl.count) pl.count[base + r] = cnt[r];
            if (pl.measured) pl.measured[base + r] = meas[r];
            if (pl.lkv) pl.lkv[base + r] = lkv[r];
            if (pl.lkw) pl.lkw[base + r] = lkw[r];
            if (pl.pv) pl.pv[base + r] = pvv[r];
            if (pl.totals) pl.totals[base + r] = rock_totals_pack(ts[r], td[r]);
        }
    }
};
// The same state for a rollout that starts from FRESH planes (Rock.__init__ values, an empty history) and hands none back:
// one 32-byte record per rock in a caller-provided scratch, touched lazily.  With the planes in local memory every check
// costs ~14 scattered 4/8-byte accesses, each its own sector (measured: 16 GB of L1/L2 traffic for 2^20 envs x 32 steps);
// a record is ONE sector in and one out, and a rock that was never checked is never read or written at all.
struct alignas(32) RockRec { double lkv, lkw, pv; int16_t cnt, meas, ts, td; };
static_assert(sizeof(RockRec) == 32, "one 32-byte sector per rock");
struct RockHeurRecords {
    RockRec* rec;                           // this env's k records
    uint32_t touched;                       // rocks whose record has been written
    uint32_t m_sample, m_dir, m_check;
    POMDP_HD void init(RockRec* mine) {     // fresh planes: tot_sample = 0, tot_dir = 0, measured = count = 0, pv = .5
        rec = mine; touched = 0u; m_sample = 0u; m_dir = 0xFFFFu; m_check = 0xFFFFu;
    }
    POMDP_HD uint32_t sample_mask(int) const { return m_sample; }
    POMDP_HD uint32_t dir_mask(int) const { return m_dir; }
    POMDP_HD uint32_t check_mask(int) const { return m_check; }
    static POMDP_HD RockRec load(const RockRec* q) {
#if defined(__CUDA_ARCH__)
        union { RockRec r; uint4 v[2]; } u;
        u.v[0] = __ldcg(reinterpret_cast<const uint4*>(q));
        u.v[1] = __ldcg(reinterpret_cast<const uint4*>(q) + 1);
        return u.r;
#else
        return *q;
#endif
    }
    static POMDP_HD void store(RockRec* q, const RockRec& r) {
#if defined(__CUDA_ARCH__)
        union { RockRec r; uint4 v[2]; } u;
        u.r = r;
        __stcg(reinterpret_cast<uint4*>(q), u.v[0]);
        __stcg(reinterpret_cast<uint4*>(q) + 1, u.v[1]);
#else
        *q = r;
#endif
    }
    // rock r was checked: rock.py:177-191 (belief side-statistics) and rock.py:566 (the transition joins the history)
    template <typename S>
    POMDP_HD void check(const RockDev& p, const RockTableHdr* hdr, S s, int32_t a, int32_t ob, int32_t obs_field, int32_t next_field) {
        const int r = a - 5;
        const uint32_t bit = 1u << r;
        RockRec c;
        if (touched & bit) c = load(rec + r);
        else { c.lkv = 1.0; c.lkw = 1.0; c.pv = .5; c.cnt = 0; c.meas = 0; c.ts = 0; c.td = 0; }
        int32_t cnt = c.cnt, meas = c.meas, ts = c.ts, td = c.td;
        rock_belief_update<S>(p, hdr, s, a, ob, cnt, meas, c.lkv, c.lkw, c.pv);
        rock_history_update(a, obs_field, next_field, ts, td);
        c.cnt = (int16_t)cnt; c.meas = (int16_t)meas; c.ts = (int16_t)ts; c.td = (int16_t)td;
        store(rec + r, c);
        touched |= bit;
        m_sample = (m_sample & ~bit) | (rock_pred_sample(ts) ? bit : 0u);
        m_dir = (m_dir & ~bit) | (rock_pred_dir(td) ? bit : 0u);
        m_check = (m_check & ~bit) | (rock_pred_check(meas, cnt, c.pv) ? bit : 0u);
    }
};
// The reference's heuristic rollout loop (rock.py:557-572 with use_heuristic=True) for one env:
//   a = choice(_generate_preferred(history)); next_ob, rw, done = step(a); history.append(Transition(...)); ob = next_ob;
//   r += rw * disc; disc *= gamma.
// next_is_reward: the transition's `next_observation` field holds the reward (the reference's own positional
// Transition(ob, action, next_ob, rw, done), rock.py:566), otherwise the next observation.
template <typename S, bool STOCH, class H>
POMDP_HD void rock_rollout_preferred1(const RockDev& p, const unsigned char* tbl, S& s, const PhiloxKey& seed, uint64_t env,
                                      uint32_t ctr0, int32_t max_steps, double gamma, bool next_is_reward, bool has_first,
                                      int32_t first_action, H& h, int32_t& prev_ob, RolloutAcc& acc) {
    typedef RockEnvT<S, STOCH> Env;
    const RockTableHdr* hdr = reinterpret_cast<const RockTableHdr*>(tbl);
    const RockLut* lut = reinterpret_cast<const RockLut*>(tbl + ROCK_LUT_OFFSET);
    acc.init(Env::is_done(s));
    for (int32_t t = 0; t < max_steps && !Env::is_done(s); ++t) {
        const uint32_t ctr = ctr0 + (uint32_t)t;
        const int32_t a = (t == 0 && has_first) ? first_action
                                                : rock_policy_preferred<S>(p, hdr, lut, s, h, draw_word(seed, env, ctr, DOMAIN_POLICY, 0));
        S s2; int32_t ob, fl; float rw;
        Env::step1(p, tbl, s, a, seed, env, ctr, s2, ob, rw, fl);
        s = s2;
        if (a >= 5 && a < (int32_t)p.n_actions)
            h.template check<S>(p, hdr, s, a, ob, prev_ob, next_is_reward ? (int32_t)rw : ob);
        prev_ob = ob;                                                                                  // rock.py:567
        acc.add(Env::reward64(rw), gamma, fl);
    }
}
// tag.py:303-316 with the heuristic: a = choice(_generate_preferred(history)); ob, rw, done = step(a); history.append(a, ob)
template <int NOPP>
POMDP_HD void tag_rollout_preferred1(const TagDev& p, const unsigned char* tbl, uint32_t& s, const PhiloxKey& seed, uint64_t env,
                                     uint32_t ctr0, int32_t max_steps, double gamma, bool has_first, int32_t first_action,
                                     int32_t& last_ob, int32_t& last_action, RolloutAcc& acc) {
    typedef TagEnvT<NOPP> Env;
    const TagTables* T = reinterpret_cast<const TagTables*>(tbl);
    acc.init(Env::is_done(s));
    for (int32_t t = 0; t < max_steps && !Env::is_done(s); ++t) {
        const uint32_t ctr = ctr0 + (uint32_t)t;
        const int32_t a = (t == 0 && has_first) ? first_action
                                                : tag_policy_preferred(T, s, last_ob, last_action, draw_word(seed, env, ctr, DOMAIN_POLICY, 0));
        uint32_t s2; int32_t ob, fl; float rw;
        if (NOPP == 1) tag_step_1opp(p, T, s, a, draw_word(seed, env, ctr, DOMAIN_STEP, 0), s2, ob, rw, fl);   // one table word (TagTables.lut)
        else Env::step1(p, tbl, s, a, seed, env, ctr, s2, ob, rw, fl);
        s = s2; last_ob = ob; last_action = a;
        acc.add((double)rw, gamma, fl);
    }
}

}  // namespace pomdp
