// pomdp_core.h -- per-env-instance transition functions over PACKED int32 state words.
//
// Everything here is a register-level `__host__ __device__` inline function: the CUDA
// kernels in pomdp_kernels.cu wrap them with the memory movement (vectorised global
// loads/stores, TMA staging of the static maps), and tests/hostsim/ compiles the very same
// functions with g++ so that the kernel LOGIC can be checked against oracle/ in a
// container without a GPU.  The host build is a test vehicle only -- the shipped library
// has no CPU path.
//
// Semantics follow d3sm0/gym_pomdp (paths below are under gym_pomdp/envs/ of the
// reference); layouts and the draw-slot contract are documented in include/pomdp_b200.h
// and DESIGN.md.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define POMDP_HD __host__ __device__ __forceinline__
#define POMDP_UNROLL _Pragma("unroll")
#else
#define POMDP_HD inline
#define POMDP_UNROLL
#endif

namespace pomdp {

enum : int32_t { FLAG_DONE = 1, FLAG_BAD_ACTION = 2, FLAG_STEPPED_DONE = 4, FLAG_BAD_STATE = 8 };
enum : uint32_t { DOMAIN_STEP = 0, DOMAIN_RESET = 1, DOMAIN_POLICY = 2, DOMAIN_SHIP = 3 };

// ------------------------------------------------------------------ Philox4x32-10 ----
struct U4 { uint32_t x, y, z, w; };

POMDP_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// The ten round keys (k + i * Weyl constant) are computed once per call on the host and reach
// the kernels as a __grid_constant__ parameter, so on the device they are constant-bank
// operands of the round's LOP3: a Philox round is 2 IMAD.WIDE + 2 LOP3 and nothing else.
struct PhiloxKey { uint32_t k0[10], k1[10]; };

POMDP_HD PhiloxKey philox_key(uint64_t seed) {
    PhiloxKey k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) { k.k0[i] = a; k.k1[i] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    return k;
}

POMDP_HD U4 philox4x32_10(U4 c, const PhiloxKey& key) {
    POMDP_UNROLL
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        U4 nx;
        nx.x = hi1 ^ c.y ^ key.k0[i];
        nx.y = lo1;
        nx.z = hi0 ^ c.w ^ key.k1[i];
        nx.w = lo0;
        c = nx;
    }
    return c;
}

// Draw-slot contract v2 (include/pomdp_b200.h): ONE Philox block holds the SAME slot of FOUR
// consecutive env instances -- word(env, slot) = philox(key = seed,
// ctr = (lo32(env >> 2), hi32(env >> 2), step, (domain << 24) | slot))[env & 3] -- so a thread
// that owns an aligned group of four envs pays one Philox call per slot, not one per env.
POMDP_HD U4 draw_quad(const PhiloxKey& seed, uint64_t group, uint32_t step, uint32_t domain, uint32_t slot) {
    U4 c;
    c.x = (uint32_t)group;
    c.y = (uint32_t)(group >> 32);
    c.z = step;
    c.w = (domain << 24) | slot;
    return philox4x32_10(c, seed);
}

POMDP_HD uint32_t word_of(const U4& r, int j) { return j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w; }

POMDP_HD uint32_t draw_word(const PhiloxKey& seed, uint64_t env, uint32_t step, uint32_t domain, uint32_t slot) {
    return word_of(draw_quad(seed, env >> 2, step, domain, slot), (int)(env & 3));
}

// Draw providers handed to the per-env functors: draw(slot) -> uint32 word.
struct LazyDraw {          // scalar paths: one Philox call per requested slot
    const PhiloxKey* key;
    uint64_t env;
    uint32_t step, domain;
    POMDP_HD uint32_t operator()(int slot) const { return draw_word(*key, env, step, domain, (uint32_t)slot); }
};
// BattleShip's fixed-time placement draws are keyed by the env itself, not by its group of four: a board is one thread's
// (or one warp's) work, so a block whose four words serve four ENVS would be computed four times over.  Instead
//   word(env, ship) = philox(key = seed, ctr = (lo32(env), hi32(env), step, DOMAIN_SHIP << 24 | ship >> 2))[ship & 3]
// -- ONE Philox call per board covers ships 0..3 (the stock game has two).
struct ShipDraw {
    const PhiloxKey* key;
    uint64_t env;
    uint32_t step;
    U4 q0;
    POMDP_HD ShipDraw(const PhiloxKey& k, uint64_t env_, uint32_t step_) : key(&k), env(env_), step(step_) {
        q0 = draw_quad(k, env_, step_, DOMAIN_SHIP, 0u);
    }
    POMDP_HD uint32_t operator()(int ship) const {
        if (ship < 4) return word_of(q0, ship);
        return word_of(draw_quad(*key, env, step, DOMAIN_SHIP, (uint32_t)(ship >> 2)), ship & 3);
    }
};
template <int N>
struct WordDraw {          // vector path: words precomputed from the group's quads
    uint32_t w[N];
    POMDP_HD uint32_t operator()(int slot) const { return w[slot]; }
};

// np.random.randint(n) / choice index under the coupling rule: floor(u * n), u = r / 2^32.
POMDP_HD uint32_t rand_below(uint32_t r, uint32_t n) { return mulhi32(r, n); }
// np.random.binomial(1, p) with T = ceil(p * 2^32)  (0 <= T <= 2^32, hence 64-bit).
POMDP_HD bool bern(uint32_t r, uint64_t T) { return (uint64_t)r < T; }

POMDP_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
POMDP_HD int popc64(uint64_t v) { return popc32((uint32_t)v) + popc32((uint32_t)(v >> 32)); }

// ------------------------------------------------------------------ geometry --------
// Moves, coord.py:101-106: 0 N(0,+1) 1 E(+1,0) 2 S(0,-1) 3 W(-1,0) 4 NULL(0,0).
POMDP_HD int move_dx(int m) { return (m == 1) - (m == 3); }
POMDP_HD int move_dy(int m) { return (m == 0) - (m == 2); }
// Grid, coord.py:58-66
POMDP_HD int grid_get_index(int x_size, int x, int y) { return x_size * y + x; }
POMDP_HD bool grid_is_inside(int x_size, int y_size, int x, int y) {
    return x >= 0 && y >= 0 && x < x_size && y < y_size;
}
POMDP_HD int iabs(int v) { return v < 0 ? -v : v; }
// Grid.euclidean_distance is np.linalg.norm(., 1): the L1 norm (coord.py:79-81).
POMDP_HD int l1_distance(int x0, int y0, int x1, int y1) { return iabs(x0 - x1) + iabs(y0 - y1); }

// TagGrid, tag.py:46-66.  29 cells: two rows of 10, then a 3x3 block at x in 5..7, y in 2..4.
POMDP_HD bool tag_is_inside(int x, int y) {
    return y >= 2 ? (x >= 5 && x < 8 && y < 5) : (x >= 0 && x < 10 && y >= 0);
}
POMDP_HD void tag_get_coord(uint32_t idx, int& x, int& y) {
    if (idx < 20) {
        y = idx >= 10;
        x = (int)idx - 10 * y;
    } else {
        const uint32_t j = idx - 20;
        const int q = (int)((j * 11u) >> 5);  // j / 3 for j < 9
        x = (int)j - 3 * q + 5;
        y = q + 2;
    }
}
POMDP_HD int tag_get_index(int x, int y) { return y < 2 ? y * 10 + x : 20 + (y - 2) * 3 + x - 5; }

// ====================================================================== RockSample ===
// Static maps of one Rock configuration, built on the host (pomdp_host.h: make_rock) and
// staged into shared memory by ONE TMA bulk copy per CTA.  Byte layout:
//   RockTableHdr (688 B)         the reference's own maps (grid, rock coordinates, sensor thresholds, legal-list order)
//   RockRes rtab[64]  (16 B)     results: 8 rows x 8 entries; entry = row + 2 * status code + truthful
//   RockLut special[4] (8 B)     NOOP (failed p_move gate), STEPPED_DONE, BAD_ACTION
//   RockLut lut[rows * n_act]    transitions, indexed by (agent cell = x | y << 4, action); only the
//                                rows of reachable cells (16 * (n-1) + n of 256) are stored and copied
// The step functor is two dependent shared-memory loads and ~20 integer instructions that are
// the same for every action class, so neither the four envs of a thread nor the 32 threads
// of a warp diverge, and there is no int->float conversion.  Shared-memory bandwidth is the
// scarce resource next to the ALU pipe (measured): the lut access is random
// over ~22 KB (bank conflicts), so its entries are kept at 8 bytes; the rtab access hits
// <= 64 distinct addresses (mostly broadcasts), so its entries can afford 16 bytes whose
// words are stored as they come.
//   lut entry   x = ceil(eff(d) * 2^32) - 1 for a check of rock a-5 at L1 distance d (rock.py:383-387,
//                   401-407); 0xFFFFFFFF for the other actions
//               y = byte 0: (bit offset of the status this action looks at) - 1, so that
//                           (state >> y) & 6 = 2 * status code          [no status: RockBits::NONE_SH1]
//                   byte 1: rtab row;  byte 2: cell ^ next cell (moves, rock.py:134-158);
//                   byte 3: 6 if the status is cleared (sample, rock.py:168) else 0
//   rtab entry  x = reward (float bits), y = obs, z = flags, w = 0x80000000 if done
struct RockTableHdr {
    int8_t grid[256];      // [x | y << 4] -> rock id written by rock.py:110-111, -1 = none
    uint8_t rock_pos[16];  // rock i -> x | y << 4   (rock.py:106)
    uint32_t thr_m1[32];   // d -> ceil(eff(d) * 2^32) - 1, eff = (1 + 2^(-d/20)) / 2
    uint8_t legal_act[32]; // position in _generate_legal's candidate order (rock.py:273-291: E, N, S, W, SAMPLE, then
                           // one check per rock i in rock order) -> the action id that entry holds
    double eff[32];        // d -> eff(d) as the Python double of rock.py:383-387 (for _compute_prob, rock.py:250-264)
};
static_assert(sizeof(RockTableHdr) == 688 && sizeof(RockTableHdr) % 16 == 0, "TMA bulk copy needs 16 B multiples");
struct alignas(8) RockLut { uint32_t x, y; };
struct alignas(16) RockRes { uint32_t x, y, z, w; };

enum : uint32_t {   // rtab rows (entry index of the row's first entry) and special lut entries
    ROCK_ROW_ZERO = 0, ROCK_ROW_EXIT = 8, ROCK_ROW_WALL = 16, ROCK_ROW_SAMPLE = 24, ROCK_ROW_DANGLING = 32,
    ROCK_ROW_CHECK = 40, ROCK_ROW_STEPPED_DONE = 48, ROCK_ROW_BAD_ACTION = 56, ROCK_RTAB_ENTRIES = 64,
    ROCK_IDX_NOOP = 0, ROCK_IDX_STEPPED_DONE = 1, ROCK_IDX_BAD_ACTION = 2, ROCK_SPECIALS = 4,
};
constexpr uint32_t ROCK_RTAB_OFFSET = sizeof(RockTableHdr);
constexpr uint32_t ROCK_LUT_OFFSET = ROCK_RTAB_OFFSET + ROCK_RTAB_ENTRIES * sizeof(RockRes);   // specials, then rows

struct RockDev {  // passed by value to the kernels
    int32_t n, k;
    int32_t stochastic;
    int32_t penal;        // rock.py:117 (-100) / rock.py:432 (0)
    uint32_t start;       // x | y << 4 of config init_pos
    uint32_t n_actions;   // 5 + k (rock.py:113)
    uint32_t lut_stride;  // lut row pitch in entries: n_actions | 1 (odd), see rock_lut_index
    uint32_t table_bytes; // bytes the TMA copy moves (header + rtab + specials + stored rows)
    uint32_t smem_bytes;  // shared-memory allocation: room for all 256 rows, so that a corrupt
                          // cell index reads garbage instead of faulting
    uint32_t gate_on;     // p_move > 0
    uint32_t gate_thr_m1; // ceil(p_move * 2^32) - 1
};

template <typename S> struct RockBits;
template <> struct RockBits<uint32_t> {
    static constexpr uint32_t DONE = 0x80000000u;
    static constexpr uint32_t NONE_SH1 = 29;   // status bits 30-31: unused / done (0 in a steppable state)
};
template <> struct RockBits<uint64_t> {
    static constexpr uint64_t DONE = 0x8000000000000000ull;
    static constexpr uint32_t NONE_SH1 = 61;
};

// shifts whose count wraps at the word size (one SHF on the GPU; upper bits of the count are ignored)
POMDP_HD uint32_t shr_wrap(uint32_t v, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(v, 0u, sh);
#else
    return v >> (sh & 31u);
#endif
}
POMDP_HD uint32_t shr_wrap(uint64_t v, uint32_t sh) { return (uint32_t)(v >> (sh & 63u)); }
POMDP_HD uint32_t shl_wrap(uint32_t v, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(0u, v, sh);
#else
    return v << (sh & 31u);
#endif
}
POMDP_HD uint64_t shl_wrap(uint64_t v, uint32_t sh) { return v << (sh & 63u); }
POMDP_HD uint32_t byte_of(uint32_t v, int b) {   // one PRMT
#if defined(__CUDA_ARCH__)
    return __byte_perm(v, 0u, 0x4440u | (uint32_t)b);
#else
    return (v >> (8 * b)) & 0xFFu;
#endif
}
POMDP_HD float bits_to_float(uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f;
    __builtin_memcpy(&f, &b, 4);
    return f;
#endif
}

// Entry of (agent cell = x | y << 4, action) in the transition LUT.  The row pitch is ODD and every row is skewed by its
// y: with the plain `cell * n_actions + action` and a 16-entry row (Rock(11,11)) the shared-memory bank of an entry
// depends on the action alone, so a batch that steps every particle with the SAME action -- what a particle filter
// does -- is a 32-way bank conflict (measured: 21.5 instead of 17.4 us per 2^22-env launch).  With an odd pitch x spreads
// the lanes over the banks, and the skew does the same for y (16 y * pitch alone is a multiple of the bank period).
POMDP_HD uint32_t rock_lut_index(const RockDev& p, uint32_t cell, uint32_t a) {
    return ROCK_SPECIALS + cell * p.lut_stride + (cell >> 4) + a;
}
constexpr uint32_t ROCK_LUT_SKEW_MAX = 15;     // entries past the last row's end that the skew can reach

// rock.py:123-194 (RockEnv.step) and rock.py:434-504 (StochasticRockEnv.step).
//   lut      -> special[0] (the rows follow at lut + ROCK_SPECIALS), rtab -> rtab[0]
//   w_gate   = draw slot 0 (p_move gate, StochasticRock only, rock.py:443)
//   w_sensor = draw slot 1 (np.random.binomial(1, eff), rock.py:404)
template <typename S, bool STOCH>
POMDP_HD void rock_step(const RockDev& p, const RockLut* __restrict__ lut, const RockRes* __restrict__ rtab, S s,
                        int32_t a, uint32_t w_gate, uint32_t w_sensor, S& s2, int32_t& ob, float& rw, int32_t& fl) {
    uint32_t idx = rock_lut_index(p, (uint32_t)s & 0xFFu, (uint32_t)a);
    if (STOCH) idx = (p.gate_on && w_gate <= p.gate_thr_m1) ? idx : (uint32_t)ROCK_IDX_NOOP;   // rock.py:443
    idx = (uint32_t)a >= p.n_actions ? (uint32_t)ROCK_IDX_BAD_ACTION : idx;                    // rock.py:125
    idx = (s & RockBits<S>::DONE) ? (uint32_t)ROCK_IDX_STEPPED_DONE : idx;                     // rock.py:126
    const RockLut e = lut[idx];
    const uint32_t code2 = shr_wrap(s, e.y) & 6u;                    // 2 * status code: 1 good, 3 bad, 0 collected / none
    const uint32_t truthful = w_sensor <= e.x ? 1u : 0u;             // rock.py:404
    const RockRes r = rtab[byte_of(e.y, 1) + code2 + truthful];
    const S clear = shl_wrap((S)byte_of(e.y, 3), e.y);               // sample: the rock's two status bits
    S ns = (s ^ (S)byte_of(e.y, 2)) & ~clear;                        // move: cell ^= delta
    ns |= (S)r.w << (8 * sizeof(S) - 32);                            // done
    s2 = ns;
    rw = bits_to_float(r.x);
    ob = (int32_t)r.y;
    fl = (int32_t)r.z;
}

// rock.py:236-241, 266-271, 78-80: status = int(sign(U(0,1) - .5)), one uniform per rock.  Only the comparison with
// one half matters, so ALL rocks (k <= 16) share ONE draw word, reset slot 0: rock i's uniform is u = r_i / 2^32 with
// r_i = rotl32(word, 30 - 2i).  The deciding (top) bit of r_i is bit 2i + 1 of the word -- sixteen different bits,
// hence independent fair coins -- and it sits exactly where the packed state keeps the HIGH bit of rock i's 2-bit
// status code (01 good, 11 bad), so the sixteen codes of a word are one LOP3: 0x55555555 | (~w & 0xAAAAAAAA).
// The tie r_i == 2^31 (u == .5, sign() gives 0 = collected) needs the word to be exactly the single bit 2i + 1.
POMDP_HD uint32_t rotl32(uint32_t v, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(v, v, sh);
#else
    sh &= 31u;
    return sh ? (v << sh) | (v >> (32u - sh)) : v;
#endif
}
POMDP_HD uint32_t rock_reset_word(uint32_t slot_word, int rock) { return rotl32(slot_word, (30u - 2u * (uint32_t)rock) & 31u); }
POMDP_HD uint32_t rock_status_code(uint32_t w) { return w > 0x80000000u ? 1u : (w < 0x80000000u ? 3u : 0u); }

// The sixteen 2-bit status codes of the reset word at once (rock i in bits 2i..2i+1).
// Equals rock_status_code(rock_reset_word(w, i)) for every w and i (brute-force-checked in tests/test_oracle_c_golden.py).
POMDP_HD uint32_t rock_reset_codes16(uint32_t w) {
    uint32_t codes = 0x55555555u | (~w & 0xAAAAAAAAu);
    if ((w & (w - 1u)) == 0u && (w & 0xAAAAAAAAu)) codes &= ~(3u * (w >> 1));   // a single bit at a deciding position: that rock ties
    return codes;
}

template <typename S>
POMDP_HD S rock_reset_from_word(const RockDev& p, uint32_t w) {
    const uint32_t m = p.k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * p.k)) - 1u);
    return (S)p.start | ((S)(rock_reset_codes16(w) & m) << 8);
}
template <typename S, class D>
POMDP_HD S rock_reset(const RockDev& p, const D& draw) { return rock_reset_from_word<S>(p, draw(0)); }
POMDP_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
POMDP_HD uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }
// Four envs of one aligned group: ONE Philox call, then per env one shift and one LOP3 -- with
// C1 = start | (0x55555555 & m) << 8 and C2 = (0xAAAAAAAA & m) << 8 (loop invariants) the packed state is
// C1 | (~(w << 8) & C2).  The tie (a word that is a single bit) is tested once for the four words and handled out
// of line: it happens with probability 2^-27 per group.
template <typename S>
POMDP_HD void rock_reset4_words(const RockDev& p, const U4& q, S out[4]) {
    const uint32_t m = p.k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * p.k)) - 1u);
    const S c1 = (S)p.start | ((S)(0x55555555u & m) << 8), c2 = (S)(0xAAAAAAAAu & m) << 8;
    out[0] = c1 | (~((S)q.x << 8) & c2);
    out[1] = c1 | (~((S)q.y << 8) & c2);
    out[2] = c1 | (~((S)q.z << 8) & c2);
    out[3] = c1 | (~((S)q.w << 8) & c2);
    const uint32_t single = umin32(umin32(q.x & (q.x - 1u), q.y & (q.y - 1u)), umin32(q.z & (q.z - 1u), q.w & (q.w - 1u)));
    if (single == 0u) {                       // some word has at most one bit set: redo the group the careful way
        out[0] = rock_reset_from_word<S>(p, q.x);
        out[1] = rock_reset_from_word<S>(p, q.y);
        out[2] = rock_reset_from_word<S>(p, q.z);
        out[3] = rock_reset_from_word<S>(p, q.w);
    }
}
template <typename S>
POMDP_HD void rock_reset4(const RockDev& p, const PhiloxKey& seed, uint64_t group, uint32_t step, S out[4]) {
    rock_reset4_words<S>(p, draw_quad(seed, group, step, DOMAIN_RESET, 0u), out);
}

// ---- uniform-legal policy: np.random.choice(env._generate_legal()) --------------------
POMDP_HD int nth_set_bit(uint32_t m, uint32_t j) {   // index of the (j+1)-th set bit of m (j < popc(m))
    // five-step binary search on masked popcounts (branch-free; CUDA's __fns is a loop)
    uint32_t pos = 0;
    POMDP_UNROLL
    for (uint32_t half = 16; half >= 1; half >>= 1) {
        const uint32_t c = (uint32_t)popc32(m & (((1u << half) - 1u) << pos));
        const bool right = j >= c;
        j -= right ? c : 0u;
        pos += right ? half : 0u;
    }
    return (int)pos;
}
// bits 0, 2, 4, ... of v gathered into the low half
POMDP_HD uint32_t compress_even(uint32_t v) {
    v &= 0x55555555u;
    v = (v | (v >> 1)) & 0x33333333u;
    v = (v | (v >> 2)) & 0x0F0F0F0Fu;
    v = (v | (v >> 4)) & 0x00FF00FFu;
    v = (v | (v >> 8)) & 0x0000FFFFu;
    return v;
}
POMDP_HD uint32_t rock_alive_bits(uint32_t s) { return compress_even(s >> 8); }           // bit i: rock i status != 0
POMDP_HD uint32_t rock_alive_bits(uint64_t s) {
    return compress_even((uint32_t)(s >> 8)) | (compress_even((uint32_t)(s >> 40)) << 16);
}
// rock.py:273-291 as a bit mask in the reference's LIST order (bit 0 E, 1 N, 2 S, 3 W, 4 SAMPLE, 5+i check of
// rock i); hdr->legal_act maps a list position to its action id.  The dangling cell of Rock(15,15)/(7,7), where the
// reference's own _generate_legal raises IndexError, offers no SAMPLE.
template <typename S>
POMDP_HD uint32_t rock_legal_list(const RockDev& p, const RockLut* __restrict__ lut, S s) {
    s &= ~RockBits<S>::DONE;                                           // the list is a function of (agent, rocks) only
    const uint32_t x = (uint32_t)s & 15u, y = ((uint32_t)s >> 4) & 15u;
    uint32_t m = 1u | ((y + 1u < (uint32_t)p.n) ? 2u : 0u) | (y > 0u ? 4u : 0u) | (x > 0u ? 8u : 0u);
    const RockLut e = lut[rock_lut_index(p, (uint32_t)s & 0xFFu, 4u)];
    m |= (shr_wrap(s, e.y) & 6u) ? 16u : 0u;                           // a rock under the agent that is not collected
    m |= (rock_alive_bits(s) & ((1u << p.k) - 1u)) << 5;
    return m;
}
template <typename S>
POMDP_HD int32_t rock_policy(const RockDev& p, const RockTableHdr* __restrict__ hdr, const RockLut* __restrict__ lut, S s,
                             uint32_t w) {
    const uint32_t m = rock_legal_list<S>(p, lut, s);
    return (int32_t)hdr->legal_act[nth_set_bit(m, rand_below(w, (uint32_t)popc32(m)))];
}

// ---- heuristic action sets: RockEnv._generate_preferred(history), rock.py:293-374 ------
// What the reference computes from the caller's History on every call -- two per-rock totals over the transitions
// whose action checked that rock -- is kept as running sums, updated once per step (rock_history_update):
//   tot_sample[i] = #(next_observation == GOOD) - #(next_observation == BAD)                       rock.py:302-309
//   tot_dir[i]    = #(next_observation == GOOD) - #(next_observation != GOOD and observation == BAD) rock.py:325-331
// (the second loop really reads `transition.observation` in its elif, rock.py:330).  The fields are taken as the
// caller's transitions hold them: a caller that builds Transition(ob, action, next_ob, rw, done) positionally -- as the
// reference's own loop does, rock.py:566 -- has the REWARD in `next_observation`; that is the caller's business and is
// selected by `next_obs_field` below, not interpreted here.  The per-rock belief side-statistics (count, measured,
// prob_valuable; rock.py:177-191) are the env's own state (rock_belief_update).
// H provides tot_sample(i), tot_dir(i), count(i), measured(i), pv(i).
// Returns a bit mask over ACTION ids -- the reference's lists are in increasing action order (NORTH 0, EAST 1, SOUTH 2,
// WEST 3, then the checks 5 + i) -- or 0 when the list came out empty and the reference falls back to _generate_legal().
// The three per-rock predicates the rule reads from the side-state.  They change only when that rock is checked, so a
// fused rollout keeps them as three bit masks in registers (RockHeurLocal) instead of re-reading five planes per rock
// per step.
POMDP_HD bool rock_pred_sample(int32_t tot_sample) { return tot_sample > 0; }                       // rock.py:310
POMDP_HD bool rock_pred_dir(int32_t tot_dir) { return tot_dir >= 0; }                               // rock.py:333
POMDP_HD bool rock_pred_check(int32_t measured, int32_t count, double pv) {                        // rock.py:368-369
    return measured < 5 && (count < 0 ? -count : count) < 2 && 0.0 < pv && pv < 1.0;
}

// H provides sample_mask(k), dir_mask(k), check_mask(k): bit i = the predicate above for rock i.
template <typename S, class H>
POMDP_HD uint32_t rock_preferred_mask(const RockDev& p, const RockTableHdr* __restrict__ hdr, S s, const H& h) {
    s &= ~RockBits<S>::DONE;
    const int x = (int)((uint32_t)s & 15u), y = (int)(((uint32_t)s >> 4) & 15u);
    const int rock = hdr->grid[(uint32_t)s & 0xFFu];                                       // rock.py:300
    const uint32_t alive = rock_alive_bits(s) & ((1u << p.k) - 1u);                        // status != 0
    // sample a rock whose checks came out good more often than bad (rock.py:301-311); a dangling grid id (Rock(15,15),
    // Rock(7,7): the reference raises IndexError at rock.py:301) counts as no rock
    if (rock >= 0 && rock < p.k && ((alive & h.sample_mask(p.k)) >> rock & 1u)) return 1u << 4;
    bool north = false, south = false, west = false, east = false;
    const uint32_t cand = alive & h.dir_mask(p.k);                                         // rock.py:322-343
    for (int i = 0; i < p.k; ++i) {
        if (!((cand >> i) & 1u)) continue;
        const int rx = hdr->rock_pos[i] & 15, ry = hdr->rock_pos[i] >> 4;
        if (ry > y) north = true;
        else if (ry < y) south = true;
        else if (rx < x) west = true;
        else if (rx > x) east = true;
    }
    if (cand == 0u) return 1u << 1;                                                        // rock.py:345-347: all_bad -> [EAST]
    uint32_t m = 0;
    if (y + 1 < p.n && north) m |= 1u << 0;                                                // rock.py:356-366
    if (east) m |= 1u << 1;
    if (y - 1 >= 0 && south) m |= 1u << 2;
    if (x - 1 >= 0 && west) m |= 1u << 3;
    m |= (alive & h.check_mask(p.k)) << 5;                                                 // rock.py:368-370
    return m;                                                                              // 0: rock.py:372-373
}
// np.random.choice(env._generate_preferred(history)) from one draw word
template <typename S, class H>
POMDP_HD int32_t rock_policy_preferred(const RockDev& p, const RockTableHdr* __restrict__ hdr, const RockLut* __restrict__ lut, S s,
                                       const H& h, uint32_t w) {
    const uint32_t m = rock_preferred_mask<S>(p, hdr, s, h);
    if (m == 0u) return rock_policy<S>(p, hdr, lut, s, w);
    return (int32_t)nth_set_bit(m, rand_below(w, (uint32_t)popc32(m)));
}
// one transition (observation field, action, next_observation field) appended to the history: the checked rock's totals
POMDP_HD void rock_history_update(int32_t a, int32_t obs_field, int32_t next_obs_field, int32_t& tot_sample, int32_t& tot_dir) {
    tot_sample += (next_obs_field == 2) - (next_obs_field == 1);
    tot_dir += next_obs_field == 2 ? 1 : (obs_field == 1 ? -1 : 0);
    (void)a;
}
// the two totals of a rock share one int32 of the check_totals plane: low half tot_sample, high half tot_dir (both signed)
POMDP_HD int32_t rock_totals_pack(int32_t tot_sample, int32_t tot_dir) {
    return (int32_t)(((uint32_t)tot_sample & 0xFFFFu) | ((uint32_t)tot_dir << 16));
}
POMDP_HD int32_t rock_totals_sample(int32_t packed) { return (int32_t)(int16_t)(packed & 0xFFFF); }
POMDP_HD int32_t rock_totals_dir(int32_t packed) { return packed >> 16; }

// ============================================================================= Tag ===
struct TagDev {
    int32_t n_opp;
    uint32_t move_on;      // move_prob > 0
    uint64_t move_T;       // ceil(move_prob * 2^32)
    uint32_t move_thr_m1;  // move_T - 1 (valid when move_on): binomial(1, move_prob) = [r <= move_thr_m1]
    uint32_t pad;
};
constexpr uint32_t TAG_DONE = 0x80000000u;
constexpr int TAG_CELLS = 29;
constexpr int TAG_MAX_OPP = 4;

POMDP_HD int tag_num_opp(uint32_t s) { return ((int32_t)(s << 1)) >> 26; }  // bits 25-30, sign-extended
POMDP_HD uint32_t tag_set_num_opp(uint32_t s, int v) { return (s & ~(63u << 25)) | (((uint32_t)v & 63u) << 25); }

// tag.py:260-280: the multiset of Moves the opponent draws from, as 2-bit codes packed
// little-end first; returns its length (2 or 4 for distinct cells).
POMDP_HD int tag_admissible(int ax, int ay, int ox, int oy, uint32_t& list) {
    int cnt = 0;
    list = 0;
#define POMDP_TAG_APPEND(cond, m) do { if (cond) { list |= (uint32_t)(m) << (2 * cnt); ++cnt; } } while (0)
    POMDP_TAG_APPEND(ox >= ax, 1);
    POMDP_TAG_APPEND(oy >= ay, 0);
    POMDP_TAG_APPEND(ox <= ax, 3);
    POMDP_TAG_APPEND(oy <= ay, 2);
    POMDP_TAG_APPEND(ox == ax && oy > ay, 0);
    POMDP_TAG_APPEND(oy == ay && ox > ax, 1);
    POMDP_TAG_APPEND(ox == ax && oy < ay, 2);
    POMDP_TAG_APPEND(oy == ay && ox < ax, 3);
#undef POMDP_TAG_APPEND
    return cnt;
}

// Static maps of the fixed 29-cell board, derived from the functions above on the host (pomdp_host.h:
// make_tag_table) and staged into shared memory by ONE TMA bulk copy per CTA, like Rock's.
//   pair[opp * 33 + agent]  (= pair[(state & 1023) + opp] for the one-opponent env)  bits 5c..5c+4 (c = 0..3): the opponent's cell if element c of the reference's move
//                           multiset (tag.py:260-280) is drawn -- the move's target, or `opp` itself when the target
//                           is off the board (tag.py:206-207);  bits 20-22: the multiset's length
//   mv[agent]               bits 5a..5a+4 (a = 0..3): the agent's cell after move a (tag.py:133-137)
// 32 x 32 pair entries so that a corrupt 5-bit cell id reads a zero entry instead of faulting, at a row pitch of 33:
// the agent's cell is observed, so the particles of a belief share it and differ in the opponent's cell -- with a
// pitch of 32 every lane of such a batch would read a different word of the SAME shared-memory bank.
#ifndef POMDP_TAG_PAIR_PITCH          // 32 reproduces the conflicting layout for the measurement in profiles/
#define POMDP_TAG_PAIR_PITCH 33
#endif
constexpr uint32_t TAG_PAIR_PITCH = POMDP_TAG_PAIR_PITCH;
POMDP_HD uint32_t tag_pair_index(uint32_t agent, uint32_t opp) { return opp * TAG_PAIR_PITCH + agent; }
//   lut[a * 32 * 33 + pair index]   the WHOLE transition of the stock one-opponent env for (agent, opp, action a), as
//                           differences to XOR into the state word:
//                           bits 0-19: (the opponent's next cell ^ opp) for each value 0..3 of the TOP TWO bits of the
//                           pick word (a two-element multiset appears as [c0, c0, c1, c1], so floor(u * len) and "top
//                           two bits" name the same element) -- all zero when the opponent cannot move (a move action,
//                           or a TAG that hits: tag.py:122-128);  bits 20-24: the agent's next cell ^ agent;
//                           bits 25-29: the observation (tag.py:219-226) ^ 31, which is never 0;  bit 30: a == TAG;
//                           bit 31: the TAG hits.  0 <=> not a valid (agent, opp, action).  tag_step_1opp reads nothing else.
constexpr uint32_t TAG_LUT_PLANE = 32 * TAG_PAIR_PITCH;
struct TagTables {
    uint32_t pair[32 * TAG_PAIR_PITCH];
    uint32_t mv[32];
    uint32_t lut[5 * TAG_LUT_PLANE];
};
static_assert(sizeof(TagTables) % 16 == 0, "TMA bulk copy needs 16 B multiples");
// pair + mv: all that the general functor (1..4 opponents), the heuristics and the queries read
constexpr uint32_t TAG_TABLES_BASE_BYTES = (uint32_t)((32 * TAG_PAIR_PITCH + 32) * sizeof(uint32_t));
static_assert(TAG_TABLES_BASE_BYTES % 16 == 0, "TMA bulk copy needs 16 B multiples");
POMDP_HD uint32_t tag_pair_entry(int agent, int opp) {
    int ax, ay, ox, oy;
    tag_get_coord((uint32_t)agent, ax, ay);
    tag_get_coord((uint32_t)opp, ox, oy);
    uint32_t list;
    const int cnt = tag_admissible(ax, ay, ox, oy, list);
    uint32_t e = (uint32_t)cnt << 20;
    for (int c = 0; c < 4; ++c) {
        int cell = opp;
        if (c < cnt) {
            const int m = (int)((list >> (2 * c)) & 3u);
            const int nx = ox + move_dx(m), ny = oy + move_dy(m);
            if (tag_is_inside(nx, ny)) cell = tag_get_index(nx, ny);
        }
        e |= (uint32_t)cell << (5 * c);
    }
    return e;
}
POMDP_HD uint32_t tag_mv_entry(int agent) {
    int ax, ay;
    tag_get_coord((uint32_t)agent, ax, ay);
    uint32_t e = 0;
    for (int a = 0; a < 4; ++a) {
        const int nx = ax + move_dx(a), ny = ay + move_dy(a);
        e |= (uint32_t)(tag_is_inside(nx, ny) ? tag_get_index(nx, ny) : agent) << (5 * a);
    }
    return e;
}
POMDP_HD uint32_t tag_lut_entry(int agent, int opp, int a) {
    const bool is_tag = a == 4, hit = is_tag && opp == agent;                     // tag.py:119-126
    const uint32_t agent2 = is_tag ? (uint32_t)agent : ((tag_mv_entry(agent) >> (5 * a)) & 31u);      // tag.py:133-137
    uint32_t cells = 0;                                                           // the opponent stays
    if (is_tag && !hit) {                                                         // tag.py:128, 201-207
        const uint32_t e = tag_pair_entry(agent, opp), len = e >> 20;             // len is 2 or 4 on this board
        for (uint32_t top2 = 0; top2 < 4; ++top2)
            cells |= (((e >> (5u * ((top2 * len) >> 2))) & 31u) ^ (uint32_t)opp) << (5u * top2);
    }
    const uint32_t ob = (!is_tag && agent2 == (uint32_t)opp) ? (uint32_t)TAG_CELLS : agent2;           // tag.py:219-226
    return cells | ((agent2 ^ (uint32_t)agent) << 20) | ((ob ^ 31u) << 25) | ((uint32_t)is_tag << 30) | ((uint32_t)hit << 31);
}
POMDP_HD void tag_build_tables(TagTables* T) {
    for (uint32_t i = 0; i < 32 * TAG_PAIR_PITCH; ++i) T->pair[i] = 0u;
    for (int opp = 0; opp < TAG_CELLS; ++opp)
        for (int agent = 0; agent < TAG_CELLS; ++agent)
            T->pair[tag_pair_index((uint32_t)agent, (uint32_t)opp)] = tag_pair_entry(agent, opp);   // never 0 for a valid pair
    for (int i = 0; i < 32; ++i) T->mv[i] = i < TAG_CELLS ? tag_mv_entry(i) : 0u;
    for (uint32_t i = 0; i < 5 * TAG_LUT_PLANE; ++i) T->lut[i] = 0u;
    for (int a = 0; a < 5; ++a)
        for (int opp = 0; opp < TAG_CELLS; ++opp)
            for (int agent = 0; agent < TAG_CELLS; ++agent)
                T->lut[(uint32_t)a * TAG_LUT_PLANE + tag_pair_index((uint32_t)agent, (uint32_t)opp)] = tag_lut_entry(agent, opp, a);
}

// Opponent j's move (tag.py:204-205) takes ONE draw word, slot j: np.random.binomial(1, move_prob) reads it whole
// (move iff w < T), np.random.choice(actions) reads its LOW half, i.e. the word tag_pick_word(w) = w << 16 under the
// usual floor(u * len) rule.  The multisets of this board have 2 or 4 elements, so 16 bits pick exactly uniformly, and
// the low half of a word is independent of "w < T" to within 2^-16 / move_prob (chi-square-tested against the
// reference's counts) -- one Philox call per four envs per opponent instead of two.
POMDP_HD uint32_t tag_pick_word(uint32_t w) { return w << 16; }

// The stock Tag-v0 (one opponent) without a branch: ONE table word holds the whole transition of (agent, opp, action);
// the draw only selects which of its four opponent cells is taken.  Same semantics as tag_step below (which handles
// 1..4 opponents).  _fast is the transition alone; the caller has ruled out done states, bad actions and bad cells.
//   w = draw slot 0: np.random.binomial(1, move_prob) (tag.py:204) and, through tag_pick_word, np.random.choice (tag.py:205)
POMDP_HD uint32_t tag_lut_word(const TagTables* __restrict__ T, uint32_t s, int32_t a) {
    const uint32_t ac = (uint32_t)a < 4u ? (uint32_t)a : 4u;                              // clamps the index; a > 4 is flagged by the caller
    return T->lut[ac * TAG_LUT_PLANE + (s & 1023u) + (TAG_PAIR_PITCH - 32u) * ((s >> 5) & 31u)];
}
// the transition of an env that is neither done nor flagged (e != 0, 0 <= a <= 4); anything else: tag_step_1opp
POMDP_HD void tag_step_1opp_fast(const TagDev& p, uint32_t e, uint32_t s, uint32_t w, uint32_t& s2, int32_t& ob, float& rw,
                                 int32_t& fl) {
    const bool alive = (int32_t)(s << 1) >= (1 << 26);                                    // num_opp > 0 (6-bit two's complement, bits 25-30)
    const bool moves = alive && p.move_on && w <= p.move_thr_m1;                          // tag.py:128, 204 (TAG and miss are in the entry)
    const uint32_t d_opp = moves ? ((e >> (5u * ((w >> 14) & 3u))) & 31u) : 0u;           // tag.py:205-207: top two bits of tag_pick_word(w)
    // num_opp lives in bits 25-30: a successful TAG subtracts one in place (6-bit wrap as tag_set_num_opp does; the borrow
    // can only reach bit 31, the done bit, which is clear on this path)
    const uint32_t ns = ((s ^ ((e >> 20) & 31u) ^ (d_opp << 5)) - ((e >> 6) & (1u << 25))) & 0x7FFFFFFFu;   // tag.py:133-137
    const bool done = (ns & (63u << 25)) == 0u;                                           // tag.py:142
    s2 = ns | (done ? TAG_DONE : 0u);
    ob = (int32_t)((~e >> 25) & 31u);                                                     // tag.py:219-226
    rw = bits_to_float(((e & (1u << 30)) ? 0xC1200000u : 0xBF800000u) ^ (e & 0x80000000u));   // -1; TAG: -10, +10 when it hits (tag.py:122-131)
    fl = done ? (int32_t)FLAG_DONE : 0;
}
POMDP_HD void tag_step_1opp(const TagDev& p, const TagTables* __restrict__ T, uint32_t s, int32_t a, uint32_t w,
                            uint32_t& s2, int32_t& ob, float& rw, int32_t& fl) {
    const uint32_t e = tag_lut_word(T, s, a);
    // the reference's asserts (tag.py:109-110, 116-117): flagged, state untouched, obs = reward = 0
    const int32_t err = (s & TAG_DONE) ? (int32_t)(FLAG_DONE | FLAG_STEPPED_DONE)
                        : ((uint32_t)a >= 5u) ? (int32_t)FLAG_BAD_ACTION
                        : (e == 0u) ? (int32_t)FLAG_BAD_STATE : 0;
    if (err) { s2 = s; ob = 0; rw = 0.f; fl = err; return; }
    tag_step_1opp_fast(p, e, s, w, s2, ob, rw, fl);
}

// tag.py:108-143 (+ move_opponent 201-207, _sample_ob 219-226).
// Draw slot j serves opponent j: np.random.binomial(1, move_prob) (tag.py:204) and np.random.choice (tag.py:205, tag_pick_word).
template <class D>
POMDP_HD void tag_step(const TagDev& p, const TagTables* __restrict__ T, uint32_t s, int32_t a, const D& draw,
                       uint32_t& s2, int32_t& ob, float& rw, int32_t& fl) {
    s2 = s; ob = 0; rw = 0.f; fl = 0;
    if (s & TAG_DONE) { fl = FLAG_DONE | FLAG_STEPPED_DONE; return; }             // tag.py:110
    if ((uint32_t)a >= 5u) { fl = FLAG_BAD_ACTION; return; }                      // tag.py:109
    uint32_t agent = s & 31u;
    bool bad = agent >= (uint32_t)TAG_CELLS;
    for (int j = 0; j < p.n_opp; ++j) bad = bad || ((s >> (5 + 5 * j)) & 31u) >= (uint32_t)TAG_CELLS;
    if (bad) { fl = FLAG_BAD_STATE; return; }                                     // tag.py:116-117
    int nopp = tag_num_opp(s);
    float reward;
    if (a == 4) {                                                                 // tag.py:119-131
        bool tagged = false;
        reward = 0.f;
        POMDP_UNROLL
        for (int j = 0; j < TAG_MAX_OPP; ++j) {
            if (j >= p.n_opp) break;
            const int sh = 5 + 5 * j;
            const uint32_t o = (s2 >> sh) & 31u;
            if (o == agent) {
                reward = 10.f;
                tagged = true;
                --nopp;
            } else if (nopp > 0) {                                                // tag.py:128 (opp is inside by construction)
                const uint32_t e = T->pair[tag_pair_index(agent, o)];
                const uint32_t w = draw(j);
                if (bern(w, p.move_T)) {                                          // tag.py:204
                    const uint32_t pick = rand_below(tag_pick_word(w), e >> 20);  // tag.py:205
                    s2 = (s2 & ~(31u << sh)) | (((e >> (5u * pick)) & 31u) << sh);  // tag.py:206-207
                }
            }
        }
        if (!tagged) reward = -10.f;
    } else {                                                                      // tag.py:133-137
        reward = -1.f;
        agent = (T->mv[agent] >> (5 * a)) & 31u;
        s2 = (s2 & ~31u) | agent;
    }
    ob = (int32_t)agent;                                                          // tag.py:219-226
    if (a < 4)
        for (int j = 0; j < p.n_opp; ++j)
            if (((s2 >> (5 + 5 * j)) & 31u) == agent) ob = TAG_CELLS;
    s2 = tag_set_num_opp(s2, nopp);
    if (nopp == 0) { s2 |= TAG_DONE; fl |= FLAG_DONE; }                           // tag.py:142
    rw = reward;
}

// tag_step for the FOUR envs of one draw group with 2..4 opponents: the opponents are walked in the outer loop, so only
// the draw block of the current opponent (slot j: one Philox call for the four envs) is live at a time.  Same
// semantics, env by env, as tag_step.
POMDP_HD void tag_step4_multi(const TagDev& p, const TagTables* __restrict__ T, const uint32_t s[4], const int32_t a[4],
                              const PhiloxKey& seed, uint64_t group, uint32_t step, uint32_t s2[4], int32_t ob[4], float rw[4],
                              int32_t fl[4]) {
    int32_t err[4], nopp[4];
    bool tagged[4];
    POMDP_UNROLL
    for (int e = 0; e < 4; ++e) {
        s2[e] = s[e];
        bool bad = (s[e] & 31u) >= (uint32_t)TAG_CELLS;
        for (int j = 0; j < p.n_opp; ++j) bad = bad || ((s[e] >> (5 + 5 * j)) & 31u) >= (uint32_t)TAG_CELLS;
        err[e] = (s[e] & TAG_DONE) ? (int32_t)(FLAG_DONE | FLAG_STEPPED_DONE)                    // tag.py:110
                 : ((uint32_t)a[e] >= 5u) ? (int32_t)FLAG_BAD_ACTION                             // tag.py:109
                 : bad ? (int32_t)FLAG_BAD_STATE : 0;                                            // tag.py:116-117
        nopp[e] = tag_num_opp(s[e]);
        tagged[e] = false;
    }
    for (int j = 0; j < p.n_opp; ++j) {                                                          // tag.py:119-131
        const bool any_tag = (a[0] == 4 && !err[0]) || (a[1] == 4 && !err[1]) || (a[2] == 4 && !err[2]) || (a[3] == 4 && !err[3]);
        if (!any_tag) break;
        const U4 qm = draw_quad(seed, group, step, DOMAIN_STEP, (uint32_t)j);
        const int sh = 5 + 5 * j;
        POMDP_UNROLL
        for (int e = 0; e < 4; ++e) {
            if (err[e] || a[e] != 4) continue;
            const uint32_t agent = s[e] & 31u;
            const uint32_t o = (s2[e] >> sh) & 31u;
            if (o == agent) {
                tagged[e] = true;
                --nopp[e];
            } else if (nopp[e] > 0) {                                                            // tag.py:128
                const uint32_t en = T->pair[tag_pair_index(agent, o)];
                if (bern(word_of(qm, e), p.move_T)) {                                            // tag.py:204
                    const uint32_t pick = rand_below(tag_pick_word(word_of(qm, e)), en >> 20);   // tag.py:205
                    s2[e] = (s2[e] & ~(31u << sh)) | (((en >> (5u * pick)) & 31u) << sh);        // tag.py:206-207
                }
            }
        }
    }
    POMDP_UNROLL
    for (int e = 0; e < 4; ++e) {
        if (err[e]) { s2[e] = s[e]; ob[e] = 0; rw[e] = 0.f; fl[e] = err[e]; continue; }
        uint32_t agent = s[e] & 31u;
        float reward;
        if (a[e] == 4) {
            reward = tagged[e] ? 10.f : -10.f;
        } else {                                                                                 // tag.py:133-137
            reward = -1.f;
            agent = (T->mv[agent] >> (5 * a[e])) & 31u;
            s2[e] = (s2[e] & ~31u) | agent;
        }
        int32_t o_ = (int32_t)agent;                                                             // tag.py:219-226
        if (a[e] < 4)
            for (int j = 0; j < p.n_opp; ++j)
                if (((s2[e] >> (5 + 5 * j)) & 31u) == agent) o_ = TAG_CELLS;
        s2[e] = tag_set_num_opp(s2[e], nopp[e]);
        fl[e] = 0;
        if (nopp[e] == 0) { s2[e] |= TAG_DONE; fl[e] = FLAG_DONE; }                              // tag.py:142
        ob[e] = o_;
        rw[e] = reward;
    }
}

// tag.py:97-102, 181-193: the reference draws 1 + n_opp cells with np.random.randint(29) (agent first, then each
// opponent); ob = _sample_ob(state, 0).  The j-th of those draws (j = 0 agent, 1 + i opponent i) is the (j % 3)-th
// base-29 DIGIT of the uniform u = w / 2^32 of reset slot j / 3:  floor(frac(u * 29^d) * 29), i.e. randint's rule
// floor(u' * 29) applied to the word w * 29^d mod 2^32 -- three cells per draw word, so the stock one-opponent env
// needs ONE Philox call per four envs.  Digit d keeps 32 - 4.86 d bits of resolution (cell probabilities within
// 29 / 2^22.3 = 6e-6 relative of 1/29 for the third digit; chi-square-tested against the reference's counts).
constexpr int TAG_DIGITS_PER_WORD = 3;
POMDP_HD uint32_t tag_reset_word(uint32_t slot_word, int digit) {
    return digit == 0 ? slot_word : (digit == 1 ? slot_word * 29u : slot_word * 841u);
}
template <class D>
POMDP_HD void tag_reset(const TagDev& p, const D& draw, uint32_t& s, int32_t& ob) {
    const uint32_t w0 = draw(0);
    const uint32_t agent = rand_below(w0, TAG_CELLS);
    s = agent;
    ob = (int32_t)agent;
    POMDP_UNROLL
    for (int j = 0; j < TAG_MAX_OPP; ++j) {
        if (j >= p.n_opp) break;
        const int idx = 1 + j;
        const uint32_t w = idx < TAG_DIGITS_PER_WORD ? w0 : draw(idx / TAG_DIGITS_PER_WORD);
        const uint32_t o = rand_below(tag_reset_word(w, idx % TAG_DIGITS_PER_WORD), TAG_CELLS);
        s |= o << (5 + 5 * j);
        if (o == agent) ob = TAG_CELLS;
    }
    s = tag_set_num_opp(s, p.n_opp);
}

// TagGrid.is_corner, tag.py:68-74
POMDP_HD bool tag_is_corner(int x, int y) {
    if (!tag_is_inside(x, y)) return false;
    return y < 2 ? (x == 0 || x == 9) : (y == 4 && (x == 5 || x == 7));
}
// TagEnv._generate_preferred(history), tag.py:231-243, as a bit mask over action ids (its lists are in increasing action
// order).  last_action < 0 stands for an empty history (all five actions).  Otherwise: TAG alone when the last
// observation was 29 (an opponent on the agent's cell) and the agent stands in a corner; else every move that stays on
// the board and does not undo the last action.  0 (the reference's `assert len(actions) > 0`) cannot happen on this
// board: every cell has at least two on-board neighbours.
POMDP_HD uint32_t tag_preferred_mask(const TagTables* __restrict__ T, uint32_t s, int32_t last_ob, int32_t last_action) {
    if (last_action < 0) return 31u;
    const uint32_t agent = s & 31u;
    int x, y;
    tag_get_coord(agent < (uint32_t)TAG_CELLS ? agent : 0u, x, y);
    if (last_ob == TAG_CELLS && tag_is_corner(x, y)) return 1u << 4;
    const uint32_t mv = T->mv[agent];
    uint32_t m = 0;
    POMDP_UNROLL
    for (int d = 0; d < 4; ++d)
        if (last_action != ((d + 2) & 3) && ((mv >> (5 * d)) & 31u) != agent) m |= 1u << d;
    return m;
}
POMDP_HD int32_t tag_policy_preferred(const TagTables* __restrict__ T, uint32_t s, int32_t last_ob, int32_t last_action, uint32_t w) {
    const uint32_t m = tag_preferred_mask(T, s, last_ob, last_action);
    if (m == 0u) return 0;
    return (int32_t)nth_set_bit(m, rand_below(w, (uint32_t)popc32(m)));
}

// =========================================================================== Tiger ===
struct TigerDev {
    uint64_t listen_G;  // floor(listen_prob * 2^32):  (u > p)  <=>  (r > G)
    double listen_prob;
};
constexpr uint32_t TIGER_DONE = 0x80000000u;

// tiger.py:72-88 (+ _compute_rw 164-172, _is_terminal 155-162, _sample_state 117-119, _sample_ob 140-149)
// ONE draw word, slot 0, serves both call sites: state_space.sample() (tiger.py:118-119, after a door was opened: the word's
// top bit) and np.random.uniform() (tiger.py:143, compared with correct_prob).  The reference draws the uniform on every
// non-terminal step but reads it only after LISTEN, and draws the sample only after an OPEN: a step never consumes both, so
// sharing the word changes no distribution, joint or marginal.
template <class D>
POMDP_HD void tiger_step(const TigerDev& p, uint32_t s, int32_t a, const D& draw,
                         uint32_t& s2, int32_t& ob, float& rw, int32_t& fl) {
    s2 = s; ob = 0; rw = 0.f; fl = 0;
    if (s & TIGER_DONE) { fl = FLAG_DONE | FLAG_STEPPED_DONE; return; }
    if ((uint32_t)a >= 3u) { fl = FLAG_BAD_ACTION; return; }
    if (s > 1u) { fl = FLAG_BAD_STATE; return; }
    uint32_t st = s & 1u;
    const bool terminal = a != 2 && (uint32_t)a == st;
    rw = a == 2 ? -1.f : (terminal ? -20.f : 10.f);
    if (terminal) {                       // tiger.py:81-83: the observation returned is the state itself
        ob = (int32_t)st;
        s2 = s | TIGER_DONE;
        fl = FLAG_DONE;
        return;
    }
    const uint32_t w = draw(0);
    if (a < 2) st = rand_below(w, 2);         // state_space.sample(), tiger.py:118-119
    const bool flip = (uint64_t)w > p.listen_G;         // p > correct_prob, tiger.py:143-148 (read only when a == LISTEN)
    ob = 2;
    if (a == 2) ob = (int32_t)(flip ? 1u - st : st);
    s2 = st;
}

// tiger.py:60-66
template <class D>
POMDP_HD void tiger_reset(const D& draw, uint32_t& s, int32_t& ob) {
    s = rand_below(draw(0), 2);
    ob = 2;
}

// ========================================================================= Network ===
constexpr int NETWORK_MAX = 30;
constexpr int NET_GROUP = 5;                                   // machines whose failures ONE draw word decides
constexpr int NET_MAX_GROUPS = NETWORK_MAX / NET_GROUP;        // 6
constexpr int NET_CODES = 243;                                 // 3^5 joint outcomes of a group
constexpr int NET_COLS = 256;                                  // alias columns (the word's low 8 bits)
// One alias column (Walker/Vose): the word's upper 24 bits against `thr` pick the column's own outcome or its alias.
// masks: bits 0-4 own "fails whatever the neighbours do" machines, 5-9 own "fails under the larger probability" machines
// (a superset of bits 0-4), bits 16-20 / 21-25 the same for the alias outcome.
struct NetAlias { uint32_t thr, masks; };
// The per-configuration tables of the step, copied from the kernel parameters to shared memory once per CTA (lookups are
// per-thread divergent, which the constant bank would serialise).
struct NetworkTables {
    NetAlias alias[NET_COLS];                                  // 2 KB
    uint32_t nbd[NET_MAX_GROUPS][32];                          // OR-linear map "down machines of group g" -> machines with a down neighbour
};
struct NetworkDev {
    int32_t n, groups;               // groups = ceil(n / 5)
    uint32_t deg3;                   // machines with more than 2 neighbours (reward 2, network.py:89-92)
    uint32_t cond_flip;              // all-ones when q < p: the larger probability then belongs to "no neighbour down"
    uint64_t p_T, q_T, ob_T;         // ceil(prob * 2^32)
    uint32_t om1, ob_any;            // the observation draw as a 32-bit compare: hit = ob_any && w <= om1 (om1 = ob_T - 1)
    double p_ob;
    uint32_t nb[NETWORK_MAX + 2];    // neighbour bit masks (network.py:144-168)
    NetworkTables t;
};
constexpr uint32_t NETWORK_DONE = 0x80000000u;
// NetworkDev travels by value as a __grid_constant__ kernel parameter next to ~200 bytes of other arguments; the
// classic parameter space is 4 KB
static_assert(sizeof(NetworkDev) <= 3584, "NetworkDev no longer fits the 4 KB kernel parameter space with the other arguments");


// tenths / 10 correctly rounded to float32 without a division: one multiply by 0.1f and one
// Newton correction through two exact FMAs.  Equal to (float)t / 10.0f for every |t| <= 2^20
// (tests/test_oracle_c_golden.py walks the whole range through the host build of this function).
POMDP_HD float tenths_to_float(int t) {
    const float x = (float)t;
    const float q = x * 0.1f;
    return fmaf(fmaf(-q, 10.0f, x), 0.1f, q);
}

// Table access for network_step_n.  NetTabPtr: plain pointers (tests/hostsim, and wherever the tables sit in generic
// memory).  NetTabSmem (device): the tables in shared memory behind ONE 32-bit base address kept in a register -- through
// a generic pointer the compiler re-derives the shared window base (S2UR + UMOV + ULEA) before every lookup.
struct NetTabPtr {
    const NetworkTables* T;
    POMDP_HD NetAlias alias(uint32_t col) const { return T->alias[col]; }
    POMDP_HD uint32_t nbd(int g, uint32_t v) const { return T->nbd[g][v]; }
};
#if defined(__CUDACC__)
struct NetTabSmem {
    uint32_t base;                                               // shared-space address of a NetworkTables
    __device__ __forceinline__ NetAlias alias(uint32_t col) const {
        NetAlias e;
        asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.thr), "=r"(e.masks) : "r"(base + col * (uint32_t)sizeof(NetAlias)));
        return e;
    }
    __device__ __forceinline__ uint32_t nbd(int g, uint32_t v) const {
        uint32_t r;
        asm("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(base + (uint32_t)(sizeof(NetAlias) * NET_COLS) + ((uint32_t)g * 32u + v) * 4u));
        return r;
    }
};
#endif

// The failures among the five machines of one group from ONE draw word (include/pomdp_b200.h, "Network draws"):
// column = w & 255; the column's own outcome if w < thr (thr = 24-bit threshold << 8, so the column bits never decide),
// its alias otherwise; cond5 = which of the five machines face the larger probability.
template <class Tab>
POMDP_HD uint32_t network_group_fail(const Tab& T, uint32_t w, uint32_t cond5) {
    const NetAlias e = T.alias(w & (NET_COLS - 1));
    const uint32_t m = w < e.thr ? e.masks : e.masks >> 16;      // bits 0-4 fail either way, 5-9 fail under the larger one
    return (m | (cond5 & (m >> NET_GROUP))) & 31u;
}

// The action and the outputs of one env (network.py:87-92, 101-112) once its failures are known.
// kClean: the caller has established that this env raises no flag (not done, action in range, no stray state bits).
// Two identities keep it short: the no-op action 2n is even, so `a & 1` alone says "reboot"; and a rebooted machine
// is up, so its observation `hit` equals the ping's `bit ^ hit ^ 1` -- one expression serves both actions.
// kNarrow: at most 15 machines (known at compile time), so both population counts of the reward fit one word.
template <bool kClean, bool kNarrow = false>
POMDP_HD void network_finish(const NetworkDev& p, uint32_t all, uint32_t na, uint32_t s, int32_t a, uint32_t fail,
                             uint32_t h, uint32_t& s2, int32_t& ob, float& rw, int32_t& fl) {
    int f = 0;
    if (!kClean) {
        f = (s & ~all & ~NETWORK_DONE) ? FLAG_BAD_STATE : 0;
        f = (uint32_t)a > na ? FLAG_BAD_ACTION : f;
        f = (s & NETWORK_DONE) ? (FLAG_DONE | FLAG_STEPPED_DONE) : f;
    }
    const bool live = kClean || f == 0;
    const bool acts = (uint32_t)a < na;
    const uint32_t m = ((uint32_t)a >> 1) & 31u;
    const uint32_t reboot = (uint32_t)a & 1u;
    const uint32_t n2 = (s & ~fail) | (((kClean || acts) ? reboot : 0u) << m);      // kClean: a <= 2n, and 2n is even
    const uint32_t o = acts ? (((n2 >> m) ^ h ^ 1u) & 1u) : 2u;
    const int up = kNarrow ? popc32((kClean ? s : (s & 0xFFFFu)) + (s & p.deg3) * 65536u)   // network.py:87-92: 1 per machine that is
                           : popc32(s) + popc32(s & p.deg3);                        // up, 2 if it has more than two neighbours
    const int tenths = 10 * up - (acts ? 1 + 24 * (int)reboot : 0);
    fl = f;
    s2 = live ? n2 : s;
    ob = live ? (int32_t)o : 0;
    rw = live ? tenths_to_float(tenths) : 0.f;
}

// network.py:71-114 for L consecutive envs of ONE draw group (L = 4: a thread's aligned
// group, lane0 = 0; L = 1: a single env, lane0 = env & 3).  Slot g < groups = the joint failure
// draw of machines 5g..5g+4, slot `groups` = the action's observation draw: one Philox call per
// slot serves all L envs.  Reward is carried as an exact integer number of tenths.
//
// The reference draws binomial(1, p or q) once per machine that is up (network.py:94-99), q when a
// neighbour is down in the OLD state.  Per machine that is one uniform u against two thresholds, i.e.
// three outcomes (u < lo: fails either way; lo <= u < hi: fails under the larger probability only;
// else: stays up), independent across machines; the 3^5 joint outcomes of a group are sampled from
// one word through an alias table, and which machines face the larger probability (`cond`) selects
// between the outcome's two masks.  A failure of a machine that is already down clears a clear bit.
// G: the number of five-machine groups when it is known at compile time (2 for the stock 10-machine Network-v0: the
// group blocks then carry no uniform branches and the compiler interleaves their Philox chains), 0 = read p.groups.
template <int L, class Tab, int G = 0>
POMDP_HD void network_step_n(const NetworkDev& p, const Tab& T, const uint32_t s[L], const int32_t a[L],
                             const PhiloxKey& seed, uint64_t group, int lane0, uint32_t step,
                             uint32_t s2[L], int32_t ob[L], float rw[L], int32_t fl[L]) {
    const int groups = G ? G : p.groups;
    const uint32_t all = (1u << p.n) - 1u;
    uint32_t fail[L], cond[L], down[L];
    POMDP_UNROLL
    for (int j = 0; j < L; ++j) { cond[j] = 0; fail[j] = 0; down[j] = ~s[j] & all; }
    POMDP_UNROLL
    for (int g = 0; g < NET_MAX_GROUPS; ++g)
        if (g < groups) {                                                           // uniform branch; network.py:81-84
            POMDP_UNROLL
            for (int j = 0; j < L; ++j) cond[j] |= T.nbd(g, (down[j] >> (NET_GROUP * g)) & 31u);
        }
    POMDP_UNROLL
    for (int j = 0; j < L; ++j) cond[j] ^= p.cond_flip;         // machines that face the larger of the two probabilities
    POMDP_UNROLL
    for (int g = 0; g < NET_MAX_GROUPS; ++g)
        if (g < groups) {                                                           // network.py:94-99
            const U4 q = draw_quad(seed, group, step, DOMAIN_STEP, (uint32_t)g);
            POMDP_UNROLL
            for (int j = 0; j < L; ++j)
                fail[j] |= network_group_fail(T, word_of(q, lane0 + j), cond[j] >> (NET_GROUP * g)) << (NET_GROUP * g);
        }
    const U4 qa = draw_quad(seed, group, step, DOMAIN_STEP, (uint32_t)groups);
    const uint32_t na = (uint32_t)(2 * p.n);
    uint32_t stray = 0, amax = 0;
    POMDP_UNROLL
    for (int j = 0; j < L; ++j) {
        stray |= s[j];
        amax = (uint32_t)a[j] > amax ? (uint32_t)a[j] : amax;
    }
    if (((stray & ~all) == 0) & (amax <= na)) {            // nothing to flag in any of the L envs: the common case
        POMDP_UNROLL
        for (int j = 0; j < L; ++j) {
            const uint32_t h = (p.ob_any && word_of(qa, lane0 + j) <= p.om1) ? 1u : 0u;
            network_finish<true, (G >= 1 && G <= 3)>(p, all, na, s[j], a[j], fail[j], h, s2[j], ob[j], rw[j], fl[j]);
        }
    } else {
        POMDP_UNROLL
        for (int j = 0; j < L; ++j) {
            const uint32_t h = (p.ob_any && word_of(qa, lane0 + j) <= p.om1) ? 1u : 0u;
            network_finish<false, (G >= 1 && G <= 3)>(p, all, na, s[j], a[j], fail[j], h, s2[j], ob[j], rw[j], fl[j]);
        }
    }
}

// ====================================================================== BattleShip ===
constexpr int SHIP_WORDS = 8;
constexpr int SHIP_MAX_CELLS = 120;
constexpr int SHIP_MAX_SHIPS = 8;
struct ShipDev {
    int32_t X, Y, max_len, n_tiles;
    uint64_t col0_lo, col0_hi;   // cells with x == 0
    uint64_t colL_lo, colL_hi;   // cells with x == X-1
    // [dir][ship index] start cells from which the reference's look-ahead stays on the board: the cell
    // length + 1 steps ahead must be inside (battleship.py:199-201), ship index 0 = the longest ship
    uint64_t inside_lo[4][SHIP_MAX_SHIPS], inside_hi[4][SHIP_MAX_SHIPS];
    // [ship index] the cells of a vertical ship whose lowest cell is cell 0: bits 0, X, 2X, ... (length - 1) X
    uint64_t vpat_lo[SHIP_MAX_SHIPS], vpat_hi[SHIP_MAX_SHIPS];
    // layout of the placement tables (ShipTableHdr below; filled by pomdp_host.h: make_ship from the same enumeration
    // that builds the table, so the kernel never reads the header)
    uint32_t tbl_n0, tbl_n_tabled, tbl_off_rec, tbl_off_second;
};

// A board of up to 128 cells as two explicit 64-bit halves (bit c = cell c = X*y + x).  Nothing on the device uses
// unsigned __int128: nvcc 12.9 lowered `r & (r >> 1)` on a 128-bit value in this code to two independent 64-bit
// shifts, losing the bit that crosses from the high to the low half -- caught by the GPU parity tests, reproduced in
// isolation, absent from g++.
struct B128 { uint64_t lo, hi; };
POMDP_HD B128 b128(uint64_t lo, uint64_t hi) { B128 r; r.lo = lo; r.hi = hi; return r; }
POMDP_HD B128 operator&(B128 a, B128 b) { return b128(a.lo & b.lo, a.hi & b.hi); }
POMDP_HD B128 operator|(B128 a, B128 b) { return b128(a.lo | b.lo, a.hi | b.hi); }
POMDP_HD B128 operator~(B128 a) { return b128(~a.lo, ~a.hi); }
// shifts by 0 <= s; anything pushed past either end is dropped, s >= 128 gives 0
POMDP_HD B128 shr(B128 v, int s) {
    if (s >= 128) return b128(0, 0);
    if (s >= 64) return b128(v.hi >> (s - 64), 0);
    return b128((v.lo >> s) | ((v.hi << 1) << (63 - s)), v.hi >> s);      // (hi << 1) << (63 - s): s = 0 stays defined
}
POMDP_HD B128 shl(B128 v, int s) {
    if (s >= 128) return b128(0, 0);
    if (s >= 64) return b128(0, v.lo << (s - 64));
    return b128(v.lo << s, (v.hi << s) | ((v.lo >> 1) >> (63 - s)));
}
POMDP_HD B128 b128_bit(int i) { return i < 64 ? b128(1ull << i, 0) : b128(0, 1ull << (i - 64)); }        // 0 <= i < 128
POMDP_HD bool b128_test(B128 v, int i) { return (((i < 64 ? v.lo : v.hi) >> (i & 63)) & 1ull) != 0; }    // 0 <= i < 128

struct ShipState {
    B128 occ, vis;
    int remaining;
    bool done;
};

POMDP_HD ShipState ship_unpack(const uint32_t w[SHIP_WORDS]) {
    ShipState st;
    st.occ = b128((uint64_t)w[0] | ((uint64_t)w[1] << 32), (uint64_t)w[2] | ((uint64_t)(w[3] & 0x00FFFFFFu) << 32));
    st.vis = b128((uint64_t)w[4] | ((uint64_t)w[5] << 32), (uint64_t)w[6] | ((uint64_t)w[7] << 32));
    st.remaining = (int)((w[3] >> 24) & 0x7Fu);
    st.done = (w[3] >> 31) != 0;
    return st;
}
POMDP_HD void ship_pack(const ShipState& st, uint32_t w[SHIP_WORDS]) {
    w[0] = (uint32_t)st.occ.lo; w[1] = (uint32_t)(st.occ.lo >> 32); w[2] = (uint32_t)st.occ.hi;
    w[3] = ((uint32_t)(st.occ.hi >> 32) & 0x00FFFFFFu) | ((uint32_t)(st.remaining & 0x7F) << 24) | (st.done ? 0x80000000u : 0u);
    w[4] = (uint32_t)st.vis.lo; w[5] = (uint32_t)(st.vis.lo >> 32); w[6] = (uint32_t)st.vis.hi; w[7] = (uint32_t)(st.vis.hi >> 32);
}

// battleship.py:91-122 on the 8 packed words (the `diagonal` writes at 111-113 are dead state).
POMDP_HD void battleship_step(const ShipDev& p, const uint32_t w[SHIP_WORDS], int32_t a,
                              uint32_t w2[SHIP_WORDS], int32_t& ob, float& rw, int32_t& fl) {
    POMDP_UNROLL
    for (int i = 0; i < SHIP_WORDS; ++i) w2[i] = w[i];
    ob = 0; rw = 0.f; fl = 0;
    if (w[3] >> 31) { fl = FLAG_DONE | FLAG_STEPPED_DONE; return; }               // battleship.py:93
    if ((uint32_t)a >= (uint32_t)p.n_tiles) { fl = FLAG_BAD_ACTION; return; }     // battleship.py:94
    const int wi = a >> 5;
    const uint32_t bit = 1u << (a & 31);
    const uint32_t ow = wi == 0 ? w[0] : wi == 1 ? w[1] : wi == 2 ? w[2] : w[3];
    const uint32_t vw = wi == 0 ? w[4] : wi == 1 ? w[5] : wi == 2 ? w[6] : w[7];
    int remaining = (int)((w[3] >> 24) & 0x7Fu);
    int reward = 0;
    if (vw & bit) {
        reward = -10;
    } else {
        reward = -1;
        if (ow & bit) { ob = 1; --remaining; }
        POMDP_UNROLL
        for (int i = 0; i < 4; ++i) if (wi == i) w2[4 + i] |= bit;
    }
    bool done = false;
    if (remaining == 0) { reward += p.n_tiles; done = true; }                     // battleship.py:118-120
    w2[3] = (w2[3] & 0x00FFFFFFu) | ((uint32_t)(remaining & 0x7F) << 24) | (done ? 0x80000000u : 0u);
    if (done) fl = FLAG_DONE;
    rw = (float)reward;
}

// Cells a new ship may not touch (battleship.py:195-211): occupied cells and every cell q
// with an occupied cell at q + {N,E,S,W,NE,SE,SW}.  q + NW is never looked at (range(8)
// over the Compass enum stops before NorthWest).  The literal form: one shift per neighbour (used by the warp
// scan and the rejection loop; the bitboard placement below factors the same set into four shifts).
POMDP_HD B128 ship_blocked(const ShipDev& p, B128 occ) {
    const B128 col0 = b128(p.col0_lo, p.col0_hi), colL = b128(p.colL_lo, p.colL_hi);
    const int X = p.X;
    B128 b = occ;
    b = b | shr(occ, X);                     // q + N occupied
    b = b | shl(occ, X);                     // q + S occupied
    b = b | (shr(occ, 1) & ~colL);           // q + E occupied (q not in the last column)
    b = b | (shl(occ, 1) & ~col0);           // q + W
    b = b | (shr(occ, X + 1) & ~colL);       // q + NE
    b = b | (shl(occ, X - 1) & ~colL);       // q + SE
    b = b | (shl(occ, X + 1) & ~col0);       // q + SW
    return b;
}

// Would the reference's collision() accept this (pos, dir, length)?  Walks length + 1
// cells and needs the cell after each inside the board (battleship.py:199-201).
POMDP_HD bool ship_candidate_ok(const ShipDev& p, B128 blocked, int pos, int dir, int length) {
    const int x = pos % p.X, y = pos / p.X;
    const int dx = move_dx(dir), dy = move_dy(dir);   // Compass 0..3 == Moves 0..3
    if (!grid_is_inside(p.X, p.Y, x + (length + 1) * dx, y + (length + 1) * dy)) return false;
    const int stride = dy * p.X + dx;
    bool ok = true;
    for (int i = 0; i <= length; ++i) ok = ok && !b128_test(blocked, pos + i * stride);
    return ok;
}

// ---- bitboard placement: every (pos, dir) candidate of one ship at once ------------------------------------------
POMDP_HD B128 ship_blocked_b(const ShipDev& p, B128 occ) {
    const B128 col0 = b128(p.col0_lo, p.col0_hi), colL = b128(p.colL_lo, p.colL_hi);
    // N, S, E, W, NE, SE, SW neighbours (not NW: battleship.py:203-211 never looks there) in four shifts: with
    // v = occ | N | S, the three eastern neighbours are v >> 1 and the two western ones (occ | S) << 1.  Bits pushed
    // past cell n_tiles - 1 are off the board and masked by the caller.
    const B128 south = shl(occ, p.X);                                     // q + S occupied
    const B128 v = occ | shr(occ, p.X) | south;                           // q, q + N, q + S
    return v | (shr(v, 1) & ~colL) | (shl(occ | south, 1) & ~col0);
}
// AND of o >> (i * st) for i = 0 .. n - 1 (1 <= n <= 16) by doubling: runs of 1, 2, 4, 8 cells, then one overlapping
// step for the rest -- floor(log2 n) + 1 shifts instead of n - 1.
POMDP_HD B128 ship_run(B128 o, int st, int n) {
    B128 r = o;
    int have = 1;
    POMDP_UNROLL
    for (int k = 0; k < 4; ++k)
        if (2 * have <= n) { r = r & shr(r, have * st); have *= 2; }
    if (have < n) r = r & shr(r, (n - have) * st);
    return r;
}
// valid[d] bit pos  <=>  collision() is False for Ship(pos, direction d, length)  (battleship.py:195-211): the
// look-ahead cell is on the board and none of the length + 1 cells pos + i*dir is blocked.
POMDP_HD void ship_valid_starts(const ShipDev& p, B128 blocked, int ship_index, int length, B128 valid[4]) {
    const B128 board = p.n_tiles >= 64 ? b128(~0ull, (1ull << (p.n_tiles - 64)) - 1ull) : b128((1ull << p.n_tiles) - 1ull, 0);
    const B128 open_ = ~blocked & board;
    // Directions 0 (+X) and 1 (+1) walk up the cell index: bit pos of the run mask = all length + 1 cells
    // pos, pos + st, ... are open.  Directions 2 (-X) and 3 (-1) walk down from pos, i.e. the same run seen from its
    // other end: shift the run mask up by length * st.  The `inside` masks keep rows from wrapping (pomdp_host.h).
    const B128 run_v = ship_run(open_, p.X, length + 1), run_h = ship_run(open_, 1, length + 1);
    const B128 r[4] = {run_v, run_h, shl(run_v, length * p.X), shl(run_h, length)};
    POMDP_UNROLL
    for (int d = 0; d < 4; ++d) valid[d] = r[d] & b128(p.inside_lo[d][ship_index], p.inside_hi[d][ship_index]);
}
POMDP_HD uint32_t b128_word(B128 v, int w) { return (uint32_t)((w < 2 ? v.lo : v.hi) >> (32 * (w & 1))); }
POMDP_HD int ship_count(const B128 valid[4]) {
    int total = 0;
    POMDP_UNROLL
    for (int d = 0; d < 4; ++d) total += popc64(valid[d].lo) + popc64(valid[d].hi);
    return total;
}
// the k-th accepted candidate in increasing c = 4 * pos + dir (k < ship_count): word, then a 5-step binary search on
// the bit position with masked popcounts, then the direction
POMDP_HD int ship_pick(const B128 valid[4], int k) {
    uint32_t m[4] = {0, 0, 0, 0};
    int base = 0;
    bool found = false;
    POMDP_UNROLL
    for (int w = 0; w < 4; ++w) {
        const uint32_t a0 = b128_word(valid[0], w), a1 = b128_word(valid[1], w), a2 = b128_word(valid[2], w),
                       a3 = b128_word(valid[3], w);
        const int c = popc32(a0) + popc32(a1) + popc32(a2) + popc32(a3);
        if (!found) {
            if (k < c) { m[0] = a0; m[1] = a1; m[2] = a2; m[3] = a3; base = 32 * w; found = true; }
            else k -= c;
        }
    }
    int lo = 0;
    POMDP_UNROLL
    for (int half = 16; half >= 1; half >>= 1) {
        const uint32_t mask = ((1u << half) - 1u) << lo;
        const int c = popc32(m[0] & mask) + popc32(m[1] & mask) + popc32(m[2] & mask) + popc32(m[3] & mask);
        if (k >= c) { k -= c; lo += half; }
    }
    int dir = 0;
    bool got = false;
    POMDP_UNROLL
    for (int d = 0; d < 4; ++d)
        if (!got && ((m[d] >> lo) & 1u)) {
            if (k == 0) { dir = d; got = true; }
            else --k;
        }
    return 4 * (base + lo) + dir;
}
// marks a placement that is known to be valid on an unshot board: one shifted cell pattern
POMDP_HD B128 ship_cells(const ShipDev& p, int ship_index, int pos, int dir, int length) {
    const B128 pat = (dir & 1) ? b128((1ull << length) - 1ull, 0) : b128(p.vpat_lo[ship_index], p.vpat_hi[ship_index]);
    const int st = (dir & 1) ? 1 : p.X;
    return shl(pat, dir < 2 ? pos : pos - (length - 1) * st);
}

// battleship.py:182-193
POMDP_HD void ship_mark(const ShipDev& p, ShipState& st, int pos, int dir, int length) {
    const int stride = move_dy(dir) * p.X + move_dx(dir);
    for (int i = 0; i < length; ++i) {
        const int c = pos + i * stride;
        if (!b128_test(st.vis, c)) ++st.remaining;
        st.occ = st.occ | b128_bit(c);
    }
}

// battleship.py:167-180 as written: rejection sampling; attempt a -> slots 2a (pos), 2a+1 (dir).
POMDP_HD bool battleship_reset_rejection(const ShipDev& p, const PhiloxKey& seed, uint64_t env, uint32_t step,
                                         ShipState& st, int max_attempts) {
    st.occ = b128(0, 0); st.vis = b128(0, 0); st.remaining = 0; st.done = false;
    int a = 0;
    for (int length = p.max_len; length >= 2; --length) {
        const B128 blocked = ship_blocked(p, st.occ);
        for (;;) {
            if (a >= max_attempts) return false;
            const uint32_t w_pos = draw_word(seed, env, step, DOMAIN_RESET, (uint32_t)(2 * a));
            const uint32_t w_dir = draw_word(seed, env, step, DOMAIN_RESET, (uint32_t)(2 * a + 1));
            ++a;
            const int pos = (int)rand_below(w_pos, (uint32_t)p.n_tiles), dir = (int)rand_below(w_dir, 4);
            if (ship_candidate_ok(p, blocked, pos, dir, length)) { ship_mark(p, st, pos, dir, length); break; }
        }
    }
    return true;
}

// battleship.py:167-180 in fixed time: ship s takes the k-th of its accepted (pos, dir) candidates, k = floor(u * count),
// u = draw slot s -- the distribution of the reference's rejection loop (uniform over the accepted set).  Places ships
// ship_from, ship_from + 1, ... on top of `occ`.  Returns false when some ship has no placement (the reference would
// loop forever).
template <class D>
POMDP_HD bool battleship_place_from(const ShipDev& p, const D& draw, int ship_from, B128& occ, int& remaining) {
    int ship = ship_from;
    for (int length = p.max_len - ship_from; length >= 2; --length, ++ship) {
        B128 valid[4];
        ship_valid_starts(p, ship_blocked_b(p, occ), ship, length, valid);
        const int total = ship_count(valid);
        if (total == 0) return false;
        const int c = ship_pick(valid, (int)rand_below(draw(ship), (uint32_t)total));
        occ = occ | ship_cells(p, ship, c >> 2, c & 3, length);
        remaining += length;
    }
    return true;
}
POMDP_HD bool battleship_reset_bitboard(const ShipDev& p, const PhiloxKey& seed, uint64_t env, uint32_t step, ShipState& st) {
    st.occ = b128(0, 0); st.vis = b128(0, 0); st.remaining = 0; st.done = false;
    return battleship_place_from(p, ShipDraw(seed, env, step), 0, st.occ, st.remaining);
}

// ---- placement tables: the accepted sets of the first two ships are static --------------------------------------
// Ship 0 is placed on an empty board, so its accepted (pos, dir) set never changes (240 of 400 candidates on 10x10,
// 20 of 100 on 5x5); ship 1's accepted set depends only on where ship 0 went.  Both are tabulated on the host
// (pomdp_host.h: make_ship_table, from the very functions above) into a caller-owned device buffer:
//   ShipTableHdr (32 B)
//   rec   [n0]   {uint16 c0, uint16 n1, uint32 off1}: ship 0's k0-th accepted candidate c0 = 4 * pos + dir (in
//                increasing order), the number of candidates ship 1 then has, and where their list starts in second[]
//   second[...]  uint16  ship 1's accepted candidates per k0, each list in increasing order
// so a reset is two dependent table reads instead of ~25 128-bit shifts and a masked-popcount search per ship:
//   ship s takes list_s[floor(u_s * len(list_s))], u_s = draw slot s -- the same candidate the bitboard scan picks.
// Ships 2.. (max_len > 3) continue with the bitboard scan.
struct ShipTableHdr { uint32_t magic, n0, n_tabled, off_rec, off_second, bytes, pad0, pad1; };
struct alignas(8) ShipRec { uint16_t c0, n1; uint32_t off1; };
constexpr uint32_t SHIP_TABLE_MAGIC = 0x53485054u;
POMDP_HD uint32_t ld_ro16(const uint16_t* p) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ldg(p);
#else
    return (uint32_t)*p;
#endif
}
POMDP_HD void ld_rec(const ShipRec* p, uint32_t& c0, uint32_t& n1, uint32_t& off1) {
#if defined(__CUDA_ARCH__)
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    c0 = v.x & 0xFFFFu; n1 = v.x >> 16; off1 = v.y;
#else
    c0 = p->c0; n1 = p->n1; off1 = p->off1;
#endif
}
// kLean: every ship comes from the tables (max_len <= 3, the stock game): no scan code is instantiated, which keeps
// the kernel small enough for full occupancy.
template <bool kLean, class D>
POMDP_HD bool battleship_reset_table(const ShipDev& p, const unsigned char* __restrict__ tbl, const D& draw, ShipState& st) {
    st.occ = b128(0, 0); st.vis = b128(0, 0); st.remaining = 0; st.done = false;
    if (p.tbl_n0 == 0) return false;
    const uint32_t k0 = rand_below(draw(0), p.tbl_n0);
    uint32_t c0, n1, off1;
    ld_rec(reinterpret_cast<const ShipRec*>(tbl + p.tbl_off_rec) + k0, c0, n1, off1);
    st.occ = ship_cells(p, 0, (int)(c0 >> 2), (int)(c0 & 3u), p.max_len);
    st.remaining = p.max_len;
    if (p.max_len < 3) return true;
    if (!kLean && p.tbl_n_tabled < 2) return battleship_place_from(p, draw, 1, st.occ, st.remaining);
    if (n1 == 0) return false;
    const uint32_t k1 = rand_below(draw(1), n1);
    const uint32_t c1 = ld_ro16(reinterpret_cast<const uint16_t*>(tbl + p.tbl_off_second) + off1 + k1);
    st.occ = st.occ | ship_cells(p, 1, (int)(c1 >> 2), (int)(c1 & 3u), p.max_len - 1);
    st.remaining += p.max_len - 1;
    if (kLean) return true;
    return battleship_place_from(p, draw, 2, st.occ, st.remaining);
}

// battleship.py:157-165: _generate_legal = the unvisited cells in increasing action order; the policy draws one
// uniformly (np.random.choice).  An exhausted board (unreachable while total_remaining > 0) yields action 0.
POMDP_HD int32_t battleship_policy(const ShipDev& p, const uint32_t w[SHIP_WORDS], uint32_t word) {
    uint32_t open_[4];
    int cnt = 0;
    POMDP_UNROLL
    for (int i = 0; i < 4; ++i) {
        const int lim = p.n_tiles - 32 * i;
        const uint32_t valid = lim >= 32 ? 0xFFFFFFFFu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
        open_[i] = ~w[4 + i] & valid;
        cnt += popc32(open_[i]);
    }
    if (cnt == 0) return 0;
    int k = (int)rand_below(word, (uint32_t)cnt);
    int32_t a = 0;
    bool found = false;
    POMDP_UNROLL
    for (int i = 0; i < 4; ++i) {
        const int c = popc32(open_[i]);
        if (!found) {
            if (k < c) { a = 32 * i + nth_set_bit(open_[i], (uint32_t)k); found = true; }
            else k -= c;
        }
    }
    return a;
}

// ============================================ observation likelihoods and legal-action masks ===
// the reference's _compute_prob(action, next_state, ob) -- the particle-reweighting step
// that follows step() in a particle filter -- as float64, the values the reference's Python floats hold (1 - p is
// formed with one rounded subtraction, as `1 - eff` is in Python); and _generate_legal as a bit mask over ACTION ids.
POMDP_HD double dsub_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
POMDP_HD double dmul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
POMDP_HD double dadd_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
// rock.py:250-264
template <typename S>
POMDP_HD double rock_obs_prob(const RockDev& p, const RockTableHdr* __restrict__ hdr, S s, int32_t a, int32_t ob) {
    if (a <= 4) return ob == 0 ? 1.0 : 0.0;
    const uint32_t r = (uint32_t)(a - 5) < (uint32_t)p.k ? (uint32_t)(a - 5) : 0u;
    const uint32_t rp = hdr->rock_pos[r];
    const double eff = hdr->eff[l1_distance((int)((uint32_t)s & 15u), (int)(((uint32_t)s >> 4) & 15u), (int)(rp & 15u), (int)(rp >> 4)) & 31];
    const uint32_t code = (uint32_t)(s >> (8 + 2 * r)) & 3u;           // 1 good, 3 bad, 0 collected
    if ((ob == 2 && code == 1u) || (ob == 1 && code == 3u)) return eff;
    return dsub_rn(1.0, eff);
}
// action-indexed legal mask (rock.py:273-291; Rock(15,15)'s two rocks at (1,2) both map to action 8)
template <typename S>
POMDP_HD uint32_t rock_legal_mask(const RockDev& p, const RockTableHdr* __restrict__ hdr, const RockLut* __restrict__ lut, S s) {
    const uint32_t lst = rock_legal_list<S>(p, lut, s);
    uint32_t m = ((lst & 1u) << 1) | ((lst >> 1) & 1u) | (lst & 0x1Cu);     // E<->N swap: list [E,N,S,W,SAMPLE] -> actions 1,0,2,3,4
    for (int i = 0; i < p.k; ++i)
        if ((lst >> (5 + i)) & 1u) m |= 1u << hdr->legal_act[5 + i];
    return m;
}
// tag.py:209-217
POMDP_HD double tag_obs_prob(const TagDev& p, uint32_t s, int32_t ob) {
    const uint32_t agent = s & 31u;
    if (ob == TAG_CELLS)
        for (int j = 0; j < p.n_opp; ++j)
            if (((s >> (5 + 5 * j)) & 31u) == agent) return 1.0;
    return ob == (int32_t)agent ? 1.0 : 0.0;
}
// tiger.py:125-138
POMDP_HD double tiger_obs_prob(double correct_prob, uint32_t s, int32_t a, int32_t ob) {
    if (a == 2 && ob != 2) return (int32_t)(s & 1u) == ob ? correct_prob : dsub_rn(1.0, correct_prob);
    if (a != 2 && ob == 2) return 1.0;
    return 0.0;
}
// network.py:43-55
POMDP_HD double network_obs_prob(const NetworkDev& p, double p_ob, uint32_t s, int32_t a, int32_t ob) {
    if ((uint32_t)a < (uint32_t)(2 * p.n)) return (int32_t)((s >> (a >> 1)) & 1u) == ob ? p_ob : dsub_rn(1.0, p_ob);
    return ob == 2 ? 1.0 : 0.0;
}
// battleship.py:80-89
POMDP_HD double battleship_obs_prob(const ShipDev& p, const uint32_t w[SHIP_WORDS], int32_t a, int32_t ob) {
    const uint32_t c = (uint32_t)a < (uint32_t)p.n_tiles ? (uint32_t)a : 0u;
    const uint32_t occ = (c >> 5) == 0 ? w[0] : (c >> 5) == 1 ? w[1] : (c >> 5) == 2 ? w[2] : w[3];
    const uint32_t vis = (c >> 5) == 0 ? w[4] : (c >> 5) == 1 ? w[5] : (c >> 5) == 2 ? w[6] : w[7];
    if (ob == 0 && ((vis >> (c & 31)) & 1u)) return 1.0;
    if (ob == 1 && ((occ >> (c & 31)) & 1u)) return 1.0;
    return ob == 0 ? 1.0 : 0.0;
}

// ========================================================================== packed results ===
// Compact result word of the *_step_packed entry points (halves the bytes a host caller has to fetch):
//   bits 0-7 obs, bits 8-15 flags, bits 16-31 the reward as a signed 16-bit count of reward UNITS -- 1 for
//   Rock / Tag / BattleShip / Tiger (their rewards are integers), 0.1 for Network (rewards are tenths).
POMDP_HD int32_t pack_result(int32_t ob, int32_t units, int32_t fl) {
    return (int32_t)(((uint32_t)ob & 0xFFu) | (((uint32_t)fl & 0xFFu) << 8) | ((uint32_t)units << 16));
}
POMDP_HD int32_t reward_units_int(float rw) { return (int32_t)rw; }
POMDP_HD int32_t reward_units_tenths(float rw) {
    const float t = rw * 10.0f;
    return (int32_t)(t < 0.f ? t - 0.5f : t + 0.5f);
}

// rock.py:177-191 (and 486-500): the per-rock belief side-statistics a check updates -- measured, count and the
// likelihood products lkv / lkw with prob_valuable = .5 lkv / (.5 lkv + .5 lkw) -- in the reference's own operation
// order with separately rounded double operations, so the values (including the NaN the reference reaches once
// both products underflow) are the reference's bit for bit.  Applies when the action is a check AND produced a
// reading (StochasticRock's failed p_move gate returns obs NULL and touches nothing, rock.py:443).
POMDP_HD double ddiv_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
template <typename S>
POMDP_HD bool rock_belief_update(const RockDev& p, const RockTableHdr* __restrict__ hdr, S s, int32_t a, int32_t ob,
                                 int32_t& count, int32_t& measured, double& lkv, double& lkw, double& pv) {
    if (a <= 4 || a >= (int32_t)p.n_actions || ob == 0) return false;
    const uint32_t rp = hdr->rock_pos[a - 5];
    const double eff = hdr->eff[l1_distance((int)((uint32_t)s & 15u), (int)(((uint32_t)s >> 4) & 15u), (int)(rp & 15u), (int)(rp >> 4)) & 31];
    const double om = dsub_rn(1.0, eff);
    ++measured;
    if (ob == 2) { ++count; lkv = dmul_rn(lkv, eff); lkw = dmul_rn(lkw, om); }
    else { --count; lkw = dmul_rn(lkw, eff); lkv = dmul_rn(lkv, om); }
    const double denom = dadd_rn(dmul_rn(.5, lkv), dmul_rn(.5, lkw));
    pv = ddiv_rn(dmul_rn(.5, lkv), denom);
    return true;
}

// ========================================================================= rollouts ===
// what a POMCP simulation does with these envs (the loops at rock.py:563-572 and
// tag.py:310-316): until done or T steps,  a = np.random.choice(env._generate_legal());  ob, rw, done = env.step(a);
// r += rw * discount;  discount *= env._discount.   Step t of a rollout that starts at counter c uses counter c + t
// for BOTH its policy draw (domain POLICY, slot 0) and the step's own draws (domain STEP), so a fused rollout is, draw
// for draw, T launches of `policy` + `step` with step_ctr = c, c+1, ...  The return is accumulated in IEEE doubles
// with separately rounded multiply and add, exactly as CPython does.
struct RolloutAcc {
    double ret, disc;
    int32_t steps, flags;
    POMDP_HD void init(bool done) { ret = 0.0; disc = 1.0; steps = 0; flags = done ? (int32_t)FLAG_DONE : 0; }
    POMDP_HD void add(double reward, double gamma, int32_t fl) {
        ret = dadd_rn(ret, dmul_rn(reward, disc));
        disc = dmul_rn(disc, gamma);
        ++steps;
        flags |= fl;
    }
};

// ================================================================ belief histogram ===
// Calls add(bin) for every count this env contributes (bins: include/pomdp_b200.h).
template <class F>
POMDP_HD void belief_bins(int kind, int p0, int p1, const uint32_t* s, F add) {
    if (kind == 0) {               // Rock: p0 = k, p1 = words
        const uint64_t v = p1 == 2 ? ((uint64_t)s[0] | ((uint64_t)s[1] << 32)) : (uint64_t)s[0];
        for (int i = 0; i < p0; ++i) if (((v >> (8 + 2 * i)) & 3u) == 1u) add(i);
        add(p0 + (int)(v & 0xFF));
    } else if (kind == 1) {        // Tag
        add((int)(s[0] & 31u));
        add(TAG_CELLS + (int)((s[0] >> 5) & 31u));
    } else if (kind == 2) {        // BattleShip: p0 = n_tiles
        for (int c = 0; c < p0; ++c) if ((s[c >> 5] >> (c & 31)) & 1u) add(c);
    } else if (kind == 3) {        // Tiger
        add((int)(s[0] & 1u));
    } else {                       // Network: p0 = n
        for (int m = 0; m < p0; ++m) if ((s[0] >> m) & 1u) add(m);
    }
}

}  // namespace pomdp
