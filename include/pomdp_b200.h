/*
 * pomdp_b200.h -- C ABI of libpomdp_b200.so: batched step()/reset() generative models of
 * d3sm0/gym_pomdp's RockSample, Tag, BattleShip, Tiger and Network environments as
 * hand-written CUDA kernels for sm_100a (NVIDIA B200).
 *
 * The reference has no FFI / plugin layer at all (100 % Python).  Its
 * boundary is the old-gym Env protocol
 *      reset() -> ob ;  step(a) -> (ob, reward, done, {"state": s})
 * implemented per env in gym_pomdp/envs/{rock,tag,battleship,tiger,network}.py.  Each
 * entry point below replaces ONE of those Python methods for a whole batch of independent
 * env instances ("particles"); the citation on each function is the reference method it
 * stands in for.  INTEGRATION.md shows the ctypes binding a maintainer of the reference
 * would add.
 *
 * Conventions (all entry points; the one exception, pomdp_step_packed_host with its pipe, takes plain HOST
 * pointers, owns staging buffers and streams inside the pipe and is synchronous -- see there)
 * ------------------------------
 *  - Plain pointers and sizes only; every array pointer is DEVICE-ACCESSIBLE memory owned by
 *    the caller: device memory (torch tensors on the Python side), or pinned host memory
 *    mapped into the device address space (cudaHostAlloc under unified addressing) -- the
 *    kernels then stream over PCIe directly ("zero-copy"; used by the single-instance Python
 *    mode and by simulate_host(zero_copy=True)).  The library never allocates, frees or keeps a
 *    pointer past the call's stream ordering.  `params` structs are HOST memory, read during
 *    the call only.  Static tables (`d_table`) should live in device memory.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are
 *    asynchronous on it.  The caller selects the device (cudaSetDevice / torch device
 *    guard) -- pointers, stream and current device must agree.
 *  - Return value: 0 on success, otherwise a cudaError_t value or one of the POMDP_E_*
 *    codes below; pomdp_last_error() gives a thread-local message.  No C++ exceptions
 *    cross the ABI.
 *  - State is `words` int32 per env, row-major [n, words] (layout per env below).
 *    `next_state` may alias `state` (in-place step).  obs / flags are int32[n], reward is
 *    float[n].
 *  - flags[i]: bit 0 POMDP_FLAG_DONE (the reference's `done`), and the batched stand-ins
 *    for the reference's hot-path asserts (rock.py:125-126, tag.py:109-110,
 *    battleship.py:93-95, tiger.py:74-75, network.py:73-74), which cannot raise per
 *    element: POMDP_FLAG_BAD_ACTION (action not in action_space), POMDP_FLAG_STEPPED_DONE
 *    (state already terminal), POMDP_FLAG_BAD_STATE (state outside the env's domain, e.g.
 *    Rock(15,15)'s dangling grid id at (12,2), where the reference raises IndexError).
 *    An env with an error flag is left unchanged with obs 0 and reward 0 (BAD_STATE on a
 *    Rock sample is treated as "no rock here").
 *  - Randomness: stateless Philox4x32-10.  The word for draw slot j of env i is
 *        philox(key = seed, ctr = (lo32(g>>2), hi32(g>>2), step_ctr, (domain<<24) | j))[g & 3]
 *    with g = global_offset + i, domain 0 for step, 1 for reset and 2 for the policy draw (3: BattleShip's placement
 *    draws, keyed per env -- see pomdp_battleship_reset): one Philox block holds
 *    the same slot of four consecutive envs, so a thread that owns an aligned group of
 *    four pays one Philox call per slot.  Results do not depend on how a batch is sharded
 *    across GPUs (shards whose global_offset is a multiple of 4 take the vector path; any
 *    other offset is still correct, on the scalar path).  Slot tables and the
 *    word->decision rules (u = r / 2^32;  binomial(1,p) = [u < p];  randint(n) =
 *    floor(u*n)) are listed per env in DESIGN.md and mirror the reference's np.random
 *    call sites.
 *  - No CPU fallback: every compute entry point launches CUDA kernels and fails with a
 *    CUDA error when no device is present.
 */
#ifndef POMDP_B200_H
#define POMDP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POMDP_ABI_VERSION 15

#define POMDP_FLAG_DONE          1
#define POMDP_FLAG_BAD_ACTION    2
#define POMDP_FLAG_STEPPED_DONE  4
#define POMDP_FLAG_BAD_STATE     8

#define POMDP_E_BADARG   (-1)   /* null pointer, negative n, unsupported configuration */
#define POMDP_E_ALIGN    (-2)   /* a pointer is not 4-byte aligned */

int pomdp_abi_version(void);
const char* pomdp_last_error(void);

/* ------------------------------------------------------------------ RockSample ---- */
/* rock.py:99-118 (RockEnv.__init__), rock.py:429-432 (StochasticRockEnv.__init__).     */
typedef struct PomdpRockParams {
    int32_t board_size;   /* key of rock.config: 2, 4, 7, 11 or 15                      */
    int32_t num_rocks;    /* must be a member of config[board_size]['size'] (rock.py:101)*/
    int32_t stochastic;   /* 0 = RockEnv, 1 = StochasticRockEnv                         */
    int32_t reserved;
    double  p_move;       /* StochasticRockEnv p_move (rock.py:429); ignored otherwise   */
} PomdpRockParams;

/* State layout: 1 word if num_rocks <= 11 else 2 (little-endian 64-bit).
 *   bits 0-3 agent x, bits 4-7 agent y, bits 8+2i..9+2i rock i status as a 2-bit two's
 *   complement code (0b11 = -1 bad, 0b00 = 0 collected, 0b01 = +1 good), top bit = done. */
int     pomdp_rock_state_words(const PomdpRockParams* params);
/* Static per-config maps that the step kernel stages into shared memory with one TMA bulk
 * copy per CTA: a 688-byte header (rock-id grid, rock coordinates, sensor thresholds and efficiencies,
 * order of the legal-action list: the reference's own tables) followed by the transition LUT indexed by (agent cell, action)
 * with an odd row pitch and a per-row skew, so that a batch stepped with ONE action does not put every lane on the same
 * shared-memory bank (12.6 KB for Rock(7,8), 25.1 KB for Rock(11,11), 42.0 KB for Rock(15,15); layout in
 * gym_pomdp_b200/csrc/pomdp_core.h: rock_lut_index).
 * The caller uploads the filled buffer to the device (16-byte aligned) and passes it as
 * `d_table`.                                                                              */
int64_t pomdp_rock_table_bytes(const PomdpRockParams* params);
int     pomdp_rock_build_table(const PomdpRockParams* params, void* host_table);
/* RockEnv.step rock.py:123-194 / StochasticRockEnv.step rock.py:434-504.
 * Draw slots: 0 = p_move gate (stochastic only), 1 = sensor Bernoulli (rock.py:404).     */
int pomdp_rock_step(const PomdpRockParams* params, const void* d_table,
                    const int32_t* state, const int32_t* action,
                    int32_t* next_state, int32_t* obs, float* reward, int32_t* flags,
                    int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                    void* stream);
/* RockEnv.reset rock.py:236-241 (+ _get_init_state 266-271, Rock.__init__ 78-80).
 * Rock i's uniform(0,1) is r_i / 2^32, r_i = rotl32(word of draw slot 0, 30 - 2 i): only its comparison
 * with one half matters, so all rocks (k <= 16) share one draw word through sixteen different deciding
 * bits.  `mask` (device, uint8[n]) may be NULL = reset all; otherwise only envs with mask[i] != 0 are
 * reset (the others keep state and get obs untouched).                                   */
int pomdp_rock_reset(const PomdpRockParams* params, const void* d_table,
                     int32_t* state, int32_t* obs, const uint8_t* mask,
                     int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                     void* stream);

/* ------------------------------------------------------------------------- Tag ---- */
/* tag.py:87-95 (TagEnv.__init__).  The board is the fixed 29-cell one (tag.py:46-66).  */
typedef struct PomdpTagParams {
    int32_t num_opponents;  /* 1..4 */
    int32_t reserved;
    double  move_prob;      /* tag.py:87 */
} PomdpTagParams;
/* State: 1 word.  bits 0-4 agent cell, bits 5+5j..9+5j opponent j's cell (0..28),
 * bits 25-30 num_opp (6-bit two's complement; the reference lets it go negative with
 * several opponents), bit 31 done.                                                       */
/* Static maps of the 29-cell board (25472 bytes: for every (agent, opponent) pair the cells the
 * opponent can reach through the move multiset of tag.py:260-280, the agent's cell after each
 * move, and -- for the stock one-opponent env -- the whole transition of every (agent, opponent,
 * action) as one word; layout in gym_pomdp_b200/csrc/pomdp_core.h: TagTables).  Filled on the host, uploaded by
 * the caller (16-byte aligned) and passed as `d_table`; the kernels stage it into shared memory
 * with one TMA bulk copy per CTA.                                                             */
int64_t pomdp_tag_table_bytes(void);
int     pomdp_tag_build_table(void* host_table);
/* TagEnv.step tag.py:108-143 (+ move_opponent 201-207, _admissable_actions 260-280,
 * _sample_ob 219-226).  Draw slot j is opponent j's word w: the move Bernoulli (tag.py:204) reads it
 * whole (w < ceil(move_prob * 2^32)), the choice among the admissible moves (tag.py:205) reads its
 * low half -- floor(((w << 16) mod 2^32) * len / 2^32); len is 2 or 4 on this board.          */
int pomdp_tag_step(const PomdpTagParams* params, const void* d_table,
                   const int32_t* state, const int32_t* action,
                   int32_t* next_state, int32_t* obs, float* reward, int32_t* flags,
                   int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                   void* stream);
/* TagEnv.reset tag.py:97-102 (+ _get_init_state 181-193): 1 + num_opponents calls of randint(29), agent
 * first.  The j-th call returns the (j % 3)-th base-29 digit of draw slot j / 3's uniform, i.e.
 * floor(u' * 29) with u' = frac(u * 29^(j % 3)): three cells per draw word.                */
int pomdp_tag_reset(const PomdpTagParams* params,
                    int32_t* state, int32_t* obs, const uint8_t* mask,
                    int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                    void* stream);

/* ------------------------------------------------------------------ BattleShip ---- */
/* battleship.py:67-75 (BattleShipEnv.__init__).                                        */
typedef struct PomdpBattleshipParams {
    int32_t x_size, y_size;   /* board_size; x_size * y_size <= 120 */
    int32_t max_len;          /* ships of length max_len .. 2 (battleship.py:74-75,171) */
    int32_t reserved;
} PomdpBattleshipParams;
/* State: 8 words per env.  Cell c = x_size*y + x (the action index, coord.py:64-66).
 *   words 0-3: occupied bit c (bits 0..119); word 3 bits 24-30 total_remaining, bit 31 done
 *   words 4-7: visited bit c.                                                            */
/* BattleShipEnv.step battleship.py:91-122.  No draws.                                   */
int pomdp_battleship_step(const PomdpBattleshipParams* params,
                          const int32_t* state, const int32_t* action,
                          int32_t* next_state, int32_t* obs, float* reward, int32_t* flags,
                          int64_t n, void* stream);
/* BattleShipEnv.reset battleship.py:131-137 (+ _get_init_state 167-180, collision 195-211,
 * mark_ship 182-193) in fixed time: all 4*n_tiles (pos, dir) candidates of a ship are tested with
 * the reference's collision rule and ship s takes the k-th accepted one in increasing
 * c = 4*pos + dir, k = floor(u * count), u = ship s's draw word -- the same distribution as the
 * reference's rejection loop (uniform over the accepted set).  These placement draws are keyed by the
 * env itself, not by its group of four (a board is one thread's work):
 *     word(env, s) = philox(key = seed, ctr = (lo32(g), hi32(g), step_ctr, 3 << 24 | s >> 2))[s & 3],  g = global_offset + i
 * (domain 3), so one Philox call covers a board's first four ships.  POMDP_FLAG_BAD_STATE is raised in
 * flags (may be NULL) when no placement exists (the reference would spin forever).
 *   pomdp_battleship_reset          one THREAD per env.  With `d_table` (device copy of the placement
 *                                   tables below) the accepted sets of the first two ships are read from
 *                                   the tables -- ship 0 meets an empty board, ship 1's set depends only on
 *                                   ship 0's placement -- and boards leave through a TMA tile store; with
 *                                   d_table == NULL (or for ships 2..) candidates are four 128-bit masks
 *                                   built with shifts (bitboard scan)
 *   pomdp_battleship_reset_warpscan one WARP per env, lanes test candidates, ballots count them
 * All produce identical boards.                                                               */
/* Placement tables (host-built, caller-owned, like Rock's): per accepted placement of ship 0 an 8-byte
 * record {candidate, count and offset of ship 1's accepted list}, then those lists (~140 KB for 10x10 /
 * max_len 3).  Fill a host buffer of pomdp_battleship_table_bytes() and upload it (16-byte aligned).   */
int64_t pomdp_battleship_table_bytes(const PomdpBattleshipParams* params);
int     pomdp_battleship_build_table(const PomdpBattleshipParams* params, void* host_table);
int pomdp_battleship_reset(const PomdpBattleshipParams* params, const void* d_table,
                           int32_t* state, int32_t* obs, int32_t* flags, const uint8_t* mask,
                           int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                           void* stream);
int pomdp_battleship_reset_warpscan(const PomdpBattleshipParams* params,
                           int32_t* state, int32_t* obs, int32_t* flags, const uint8_t* mask,
                           int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                           void* stream);
/* Same method, thread per env, literally the reference's rejection loop: attempt a uses
 * slot 2a = randint(n_tiles), 2a+1 = randint(4).  Exists so that reset can be checked
 * element-wise against the reference under a coupled draw stream.                        */
int pomdp_battleship_reset_rejection(const PomdpBattleshipParams* params,
                           int32_t* state, int32_t* obs, int32_t* flags, const uint8_t* mask,
                           int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                           void* stream);

/* ----------------------------------------------------------------------- Tiger ---- */
typedef struct PomdpTigerParams {
    double listen_prob;  /* tiger.py:141: _sample_ob's default .85 (self.correct_prob is unused there) */
} PomdpTigerParams;
/* State: 1 word, bit 0 = tiger door, bit 31 done.
 * TigerEnv.step tiger.py:72-88.  ONE draw word (slot 0) serves the state resample after an OPEN (its top bit) and the
 * uniform read after a LISTEN (the reference draws the uniform on every step but reads it only then): a step never
 * consumes both. */
int pomdp_tiger_step(const PomdpTigerParams* params,
                     const int32_t* state, const int32_t* action,
                     int32_t* next_state, int32_t* obs, float* reward, int32_t* flags,
                     int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                     void* stream);
/* TigerEnv.reset tiger.py:60-66.  Slot 0 = state.  obs = 2 (NULL).                       */
int pomdp_tiger_reset(const PomdpTigerParams* params,
                      int32_t* state, int32_t* obs, const uint8_t* mask,
                      int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                      void* stream);

/* --------------------------------------------------------------------- Network ---- */
/* network.py:27-38.                                                                    */
typedef struct PomdpNetworkParams {
    int32_t n_machines;     /* <= 30; 3-legs needs n >= 4 and n % 3 == 1 (network.py:155)*/
    int32_t problem_type;   /* 3 = make_3legs_neighbours, else make_ring_neighbours      */
    double  p, q, p_ob;     /* .1, .33, .95 (network.py:28-29, 57-59)                    */
} PomdpNetworkParams;
/* State: 1 word, bit m = machine m is up; bit 31 done (never set: network.py never ends).
 * NetworkEnv.step network.py:71-114.  reward is float32 of (tenths / 10).
 *
 * Network draws.  The reference draws binomial(1, p) -- or binomial(1, q) when a neighbour is down -- once per
 * machine (network.py:94-99).  For one machine that is one uniform u against the two thresholds T_p = ceil(p 2^32),
 * T_q = ceil(q 2^32), i.e. three outcomes ("digits"): 0 = u < lo (fails either way), 1 = lo <= u < hi (fails only
 * under the larger probability), 2 = stays up, with lo = min(T_p, T_q), hi = max.  Machines are independent, so the
 * 3^5 = 243 joint outcomes of FIVE machines are drawn from ONE 32-bit word through an alias table:
 *   slot g (g < G = ceil(n / 5)) decides machines 5g .. 5g+4;  slot G is the action's observation draw ([w < T_ob]).
 *   column = w & 255;  outcome = column if w < thr24[column] << 8 (or the column is its own alias) else alias[column];
 *   outcome k = sum d_i 3^i, d_i = digit of machine 5g + i.
 * The table is built in integer arithmetic only (every implementation produces the same 256 columns):
 *   counts c = (lo, hi - lo, 2^32 - hi);  weight W[k] = chained product a <- (a * c[d_i]) >> 32 from a = 2^32;
 *   V[k] = 256 W[k] (0 for k >= 243), S = sum W;  Vose's pairing with the small (V < S) and the large columns listed
 *   in ascending k and paired from the END of both lists: thr24[s] = floor(V[s] 2^24 / S), alias[s] = l,
 *   V[l] -= S - V[s], l re-listed as small or large.
 * The outcome probabilities differ from the exact products by < 2^-24 (each of the 256 columns' 24-bit thresholds is
 * off by < 2^-32 of probability mass); 7.5e-9 at the reference's p = .1, q = .33 -- computed in exact rational
 * arithmetic by tests/test_network_draws.py.  3 Philox calls per four 10-machine envs instead of 11.             */
int pomdp_network_step(const PomdpNetworkParams* params,
                       const int32_t* state, const int32_t* action,
                       int32_t* next_state, int32_t* obs, float* reward, int32_t* flags,
                       int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                       void* stream);
/* NetworkEnv.reset network.py:61-69: all machines up, obs = 0 (OFF).                     */
int pomdp_network_reset(const PomdpNetworkParams* params,
                        int32_t* state, int32_t* obs, const uint8_t* mask,
                        int64_t n, void* stream);

/* ----------------------------------------------------- step with a compact result --- */
/* Same transition as pomdp_E_step (same draws, same next_state), but obs / reward / flags leave as
 * ONE int32 stream:   result[i] = obs | flags << 8 | reward_units << 16
 * with reward_units a signed 16-bit count of reward units: 1 for Rock, Tag and Tiger (integer
 * rewards), 0.1 for Network (rewards are tenths: s - 0.1, s - 2.5).  16 instead of 24 bytes per
 * env-step of device traffic, and 8 instead of 16 bytes per env for a host caller to fetch over
 * PCIe.  (BattleShip's 32-byte boards dominate its traffic; it has no packed variant.)          */
#define POMDP_RESULT_OBS(r)    ((int32_t)((uint32_t)(r) & 0xFFu))
#define POMDP_RESULT_FLAGS(r)  ((int32_t)(((uint32_t)(r) >> 8) & 0xFFu))
#define POMDP_RESULT_UNITS(r)  ((int32_t)(r) >> 16)
int pomdp_rock_step_packed(const PomdpRockParams* params, const void* d_table,
                           const int32_t* state, const int32_t* action, int32_t* next_state, int32_t* result,
                           int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_tag_step_packed(const PomdpTagParams* params, const void* d_table,
                          const int32_t* state, const int32_t* action, int32_t* next_state, int32_t* result,
                          int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_tiger_step_packed(const PomdpTigerParams* params,
                            const int32_t* state, const int32_t* action, int32_t* next_state, int32_t* result,
                            int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_network_step_packed(const PomdpNetworkParams* params,
                              const int32_t* state, const int32_t* action, int32_t* next_state, int32_t* result,
                              int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);

/* --------------------------------------------- step on HOST buffers (one call, pipelined) --- */
/* What a numpy-holding caller of the reference does: its (state, action) arrays live in host memory and
 * it wants (next_state, result) back in host memory (the loop around env.step at rock.py:563-572 with the
 * arrays of a whole particle set).  pomdp_step_packed_host is pomdp_E_step_packed for HOST pointers: the
 * batch is cut into chunks, and for every chunk the H2D copies, the step kernel and the D2H copies are
 * queued on the pipe's three streams (one per engine: copies in, kernels, copies out; events order a
 * chunk's three stages and guard the reuse of its staging slot), so the copies of neighbouring chunks run
 * on both PCIe directions while the kernels (microseconds) hide under them.  Returns when every result
 * has landed.  Results are identical to one pomdp_E_step_packed call over the whole batch (Philox is keyed
 * by the global index).
 *
 *  - pipe: device staging buffers ((2 * state_words + 2) * chunk_envs * 4 bytes per slot), three streams
 *    and three events per slot, created once and reused; chunk_envs must be a positive multiple of 4,
 *    1 <= n_slots <= 8.  A pipe belongs to the device that was current at creation -- the call switches to
 *    it and restores the caller's current device on return -- and is not thread-safe.
 *  - the host buffers must be HOST-COMPLETE when the call is made: it runs on the pipe's own non-blocking
 *    streams and does not order itself after work the caller has queued on other streams (an asynchronous
 *    copy into h_state, a table upload still in flight: synchronize those first).
 *  - kind: POMDP_KIND_ROCK / TAG / TIGER / NETWORK; params points at the matching PomdpEParams; d_table as
 *    for pomdp_E_step_packed (NULL for Tiger / Network), fully built before the call.
 *  - host pointers should be page-locked (cudaHostAlloc / cudaHostRegister): pageable memory is still
 *    correct, but its copies are staged by the driver and do not overlap.                               */
int pomdp_host_pipe_create(int state_words, int64_t chunk_envs, int n_slots, void** pipe_out);
int pomdp_host_pipe_destroy(void* pipe);
int pomdp_step_packed_host(void* pipe, int kind, const void* params, const void* d_table,
                           const int32_t* h_state, const int32_t* h_action, int32_t* h_next_state, int32_t* h_result,
                           int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr);

/* ------------------------------------------- uniform-legal policy and fused rollouts --- */
/* What a POMCP / Monte-Carlo caller does with these envs between two tree nodes (the loops at
 * rock.py:563-572 and tag.py:310-316):
 *
 *     while not done and t < max_steps:
 *         a = np.random.choice(env._generate_legal())         # pomdp_E_policy
 *         ob, rw, done, _ = env.step(a)                       # pomdp_E_step
 *         r += rw * discount;  discount *= gamma;  t += 1
 *
 * (rock.py:563 draws from _generate_preferred(history), which IS _generate_legal() unless the env was built
 * with use_heuristic=True -- rock.py:293-295; the history-dependent heuristic is not batched.)
 *
 * pomdp_E_policy  : action[i] = legal_i[floor(u * len(legal_i))], legal_i = the reference's
 *                   _generate_legal list IN ITS ORDER (rock.py:273-291; tag.py:228-229;
 *                   battleship.py:157-165; tiger.py:111-112; network.py:129-130), u from draw
 *                   (domain 2 = POLICY, slot 0) of `step_ctr`.
 * pomdp_E_rollout : the whole loop in ONE kernel, states in registers.  `first_action` (int32[n], may be NULL):
 *                   when given, step 0 takes the caller's action instead of a policy draw -- the Monte-Carlo
 *                   estimate of Q(s, a) a POMCP simulation needs (act, then roll out).  Step t uses counter
 *                   step_ctr + t for its policy draw and for the step's own draws, so it equals,
 *                   draw for draw, max_steps launches of pomdp_E_policy + pomdp_E_step with
 *                   step_ctr, step_ctr+1, ...  Outputs per env: final_state (may be NULL, may alias
 *                   state), ret = sum_t rw_t * discount^t as float64 accumulated exactly like the
 *                   Python loop (separately rounded multiply and add; Network's reward is the exact
 *                   double tenths/10.0), steps taken, flags = OR of the steps' flags (bit 0 = the
 *                   rollout ended in a terminal state).  An env that is already terminal takes 0 steps.
 * Rock(15,15)/(7,7): the dangling cell, where the reference's own _generate_legal raises
 * IndexError, offers no SAMPLE action.                                                         */
int pomdp_rock_policy(const PomdpRockParams* params, const void* d_table,
                      const int32_t* state, int32_t* action,
                      int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_rock_rollout(const PomdpRockParams* params, const void* d_table,
                       const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret, int32_t* steps, int32_t* flags,
                       int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                       int32_t max_steps, double discount, void* stream);
int pomdp_tag_policy(const PomdpTagParams* params, const void* d_table, const int32_t* state, int32_t* action,
                     int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_tag_rollout(const PomdpTagParams* params, const void* d_table,
                      const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret, int32_t* steps, int32_t* flags,
                      int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                      int32_t max_steps, double discount, void* stream);
int pomdp_battleship_policy(const PomdpBattleshipParams* params, const int32_t* state, int32_t* action,
                            int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_battleship_rollout(const PomdpBattleshipParams* params,
                             const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret, int32_t* steps, int32_t* flags,
                             int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                             int32_t max_steps, double discount, void* stream);
int pomdp_tiger_policy(const PomdpTigerParams* params, const int32_t* state, int32_t* action,
                       int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_tiger_rollout(const PomdpTigerParams* params,
                        const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret, int32_t* steps, int32_t* flags,
                        int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                        int32_t max_steps, double discount, void* stream);
int pomdp_network_policy(const PomdpNetworkParams* params, const int32_t* state, int32_t* action,
                         int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_network_rollout(const PomdpNetworkParams* params,
                          const int32_t* state, const int32_t* first_action, int32_t* final_state, double* ret, int32_t* steps, int32_t* flags,
                          int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                          int32_t max_steps, double discount, void* stream);

/* --------------------------------- observation likelihoods and legal-action masks --- */
/* what a particle filter / POMCP node does right after step().
 * pomdp_E_obs_prob  : prob[i] = env._compute_prob(action[i], next_state[i], obs[i]) as float64, the
 *                     value the reference's Python float holds (rock.py:250-264; tag.py:209-217;
 *                     battleship.py:80-89; tiger.py:125-138 with its `correct_prob` argument;
 *                     network.py:43-55).  `next_state` is the packed post-step state.
 * pomdp_E_legal_mask: env._generate_legal() as a bit mask over action ids, ceil(n_actions/32) uint32
 *                     words per env, row-major (rock.py:273-291; tag.py:228-229; battleship.py:157-165
 *                     = ceil(n_tiles/32) words, bit c = cell c unvisited; tiger.py:111-112; network.py:129-130).      */
int pomdp_rock_obs_prob(const PomdpRockParams* params, const void* d_table,
                        const int32_t* next_state, const int32_t* action, const int32_t* obs, double* prob,
                        int64_t n, void* stream);
int pomdp_rock_legal_mask(const PomdpRockParams* params, const void* d_table,
                          const int32_t* state, uint32_t* mask, int64_t n, void* stream);
/* RockEnv._generate_legal() as the reference's LIST (rock.py:273-291): bit b of list[i] = the b-th candidate of the
 * list's fixed order is present -- b = 0..4: EAST, NORTH, SOUTH, WEST, SAMPLE; b = 5 + r: the check the reference
 * appends for rock r, i.e. action grid[rock_pos[r]] + 5 (Rock(15,15)'s two rocks at (1,2) both give action 8, twice
 * in the list).  The header of the static table (`legal_act`, byte 400..431) maps b to its action id.          */
int pomdp_rock_legal_list(const PomdpRockParams* params, const void* d_table, const int32_t* state, uint32_t* list,
                          int64_t n, void* stream);
int pomdp_tag_obs_prob(const PomdpTagParams* params,
                       const int32_t* next_state, const int32_t* action, const int32_t* obs, double* prob,
                       int64_t n, void* stream);
int pomdp_tag_legal_mask(const PomdpTagParams* params, const int32_t* state, uint32_t* mask, int64_t n, void* stream);
int pomdp_battleship_obs_prob(const PomdpBattleshipParams* params,
                              const int32_t* next_state, const int32_t* action, const int32_t* obs, double* prob,
                              int64_t n, void* stream);
int pomdp_battleship_legal_mask(const PomdpBattleshipParams* params, const int32_t* state, uint32_t* mask,
                                int64_t n, void* stream);
int pomdp_tiger_obs_prob(const PomdpTigerParams* params,
                         const int32_t* next_state, const int32_t* action, const int32_t* obs, double* prob,
                         int64_t n, double correct_prob, void* stream);
int pomdp_tiger_legal_mask(const PomdpTigerParams* params, const int32_t* state, uint32_t* mask, int64_t n, void* stream);
int pomdp_network_obs_prob(const PomdpNetworkParams* params,
                           const int32_t* next_state, const int32_t* action, const int32_t* obs, double* prob,
                           int64_t n, void* stream);
int pomdp_network_legal_mask(const PomdpNetworkParams* params, const int32_t* state, uint32_t* mask, int64_t n, void* stream);

/* ----------------------------------------------- RockSample belief side-statistics --- */
/* rock.py:78-86, 177-191: every check of rock r updates that rock's `measured`, `count`, the
 * likelihood products `lkv` / `lkw` and `prob_valuable` = .5 lkv / (.5 lkv + .5 lkw); they are read
 * by _generate_preferred (rock.py:368) and _select_target (rock.py:394).  Batched: five arrays
 * shaped [n, num_rocks] (count, measured int32; lkv, lkw, prob_valuable float64; fresh values
 * 0, 0, 1, 1, .5 as in Rock.__init__), updated in place for every env whose `action` was a check
 * that produced a reading (obs != 0), from the post-step state.  Same double operations in the same
 * order as the reference, so the values -- including the NaN the reference itself reaches when both
 * products underflow -- are bit-identical.                                                     */
int pomdp_rock_belief_update(const PomdpRockParams* params, const void* d_table,
                             const int32_t* next_state, const int32_t* action, const int32_t* obs,
                             int32_t* count, int32_t* measured, double* lkv, double* lkw, double* prob_valuable,
                             int64_t n, void* stream);

/* ------------------------------------------------- heuristic action sets and rollouts --- */
/* RockEnv._generate_preferred(history) rock.py:293-374 (with use_heuristic=True) and the rollout loop that draws from
 * it, rock.py:557-572; TagEnv._generate_preferred(history) tag.py:231-243.
 *
 * RockSample.  The reference recomputes, from the caller's History on every call, two totals per rock over the
 * transitions whose action checked that rock (rock.py:302-309 and 325-331); here they are running sums in one int32
 * plane `check_totals[n, num_rocks]` (low 16 bits: #GOOD - #BAD of `next_observation`; high 16 bits:
 * #(next_observation == GOOD) - #(next_observation != GOOD and observation == BAD), both signed), advanced by
 * pomdp_rock_history_update with the two fields exactly as the caller's Transition holds them.  The belief
 * side-statistics it also reads (count, measured, prob_valuable) are the planes of pomdp_rock_belief_update.
 * Any plane pointer may be NULL = its fresh value (Rock.__init__ rock.py:78-86; an empty history).
 *   pomdp_rock_preferred_mask    out[i] = the preferred set as a bit mask over action ids (the reference's lists are in
 *                                increasing action order), 0 = the list came out empty and the reference returns
 *                                _generate_legal() (rock.py:372-373).  A dangling grid id under the agent
 *                                (Rock(15,15), Rock(7,7): IndexError at rock.py:301) counts as "no rock".
 *   pomdp_rock_policy_preferred  action[i] = np.random.choice(that list) (or of _generate_legal()), u from draw slot 0 of
 *                                the POLICY domain at step_ctr, like pomdp_rock_policy.
 *   pomdp_rock_rollout_preferred the whole loop for up to max_steps steps in one kernel (one thread per env, planes in
 *                                local memory): planes that are passed are read at the start and updated in place.
 *                                next_is_reward != 0: the transition's `next_observation` field holds the REWARD, which is
 *                                what the reference's own loop stores (positional Transition(ob, action, next_ob, rw,
 *                                done) against the field order (observation, action, reward, next_observation, done),
 *                                rock.py:525-530, 566); 0: it holds the next observation.                      */
typedef struct PomdpRockHeuristicPlanes {
    int32_t* count;          /* [n, num_rocks]  rock.py:83 */
    int32_t* measured;       /* [n, num_rocks]  rock.py:84 */
    double*  lkv;            /* [n, num_rocks]  rock.py:86 */
    double*  lkw;            /* [n, num_rocks]  rock.py:85 */
    double*  prob_valuable;  /* [n, num_rocks]  rock.py:87 */
    int32_t* check_totals;   /* [n, num_rocks]  see above */
    int32_t* prev_obs;       /* [n] the observation the next transition records as `observation` (reset: 0) */
    void*    scratch;        /* optional, 32-byte aligned, n * num_rocks * 32 bytes, contents irrelevant: when none of the six
                                [n, num_rocks] planes above is passed (fresh planes in, none out) and max_steps <= 32767, the
                                rollout keeps its per-rock side-state here as one lazily touched 32-byte record per rock
                                instead of in thread-local arrays (one sector in and out per check instead of ~14)          */
} PomdpRockHeuristicPlanes;
int pomdp_rock_history_update(const PomdpRockParams* params, const int32_t* observation_field, const int32_t* action,
                              const int32_t* next_observation_field, int32_t* check_totals, int64_t n, void* stream);
int pomdp_rock_preferred_mask(const PomdpRockParams* params, const void* d_table, const int32_t* state,
                              const int32_t* count, const int32_t* measured, const double* prob_valuable,
                              const int32_t* check_totals, uint32_t* mask, int64_t n, void* stream);
int pomdp_rock_policy_preferred(const PomdpRockParams* params, const void* d_table, const int32_t* state,
                                const int32_t* count, const int32_t* measured, const double* prob_valuable,
                                const int32_t* check_totals, int32_t* action, int64_t n, int64_t global_offset,
                                uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_rock_rollout_preferred(const PomdpRockParams* params, const void* d_table, const int32_t* state,
                                 const int32_t* first_action, const PomdpRockHeuristicPlanes* planes /* may be NULL */,
                                 int32_t* final_state, double* ret, int32_t* steps, int32_t* flags,
                                 int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                                 int32_t max_steps, double discount, int32_t next_is_reward, void* stream);
/* Tag: the set depends on the last (observation, action) of the history only.  last_action[i] < 0 (or a NULL array) = an
 * empty history.  The rollout updates last_obs / last_action in place when they are passed.                    */
int pomdp_tag_preferred_mask(const PomdpTagParams* params, const void* d_table, const int32_t* state,
                             const int32_t* last_obs, const int32_t* last_action, uint32_t* mask, int64_t n, void* stream);
int pomdp_tag_policy_preferred(const PomdpTagParams* params, const void* d_table, const int32_t* state,
                               const int32_t* last_obs, const int32_t* last_action, int32_t* action,
                               int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr, void* stream);
int pomdp_tag_rollout_preferred(const PomdpTagParams* params, const void* d_table, const int32_t* state,
                                int32_t* last_obs, int32_t* last_action, const int32_t* first_action,
                                int32_t* final_state, double* ret, int32_t* steps, int32_t* flags,
                                int64_t n, int64_t global_offset, uint64_t seed, uint32_t step_ctr,
                                int32_t max_steps, double discount, void* stream);

/* ------------------------------------------------------------ Grid / Coord helpers --- */
/* coord.py:7-114 and tag.py:36-66 as batched device functions (bit-exact integer work).
 * Coordinates travel as int32 pairs (x, y), i.e. arrays shaped [n, 2].
 *   POMDP_COORD_GET_INDEX : a = coords[n,2]            -> out[n]   = x_size*y + x  (coord.py:58-59)
 *   POMDP_COORD_GET_COORD : a = idx[n]                 -> out[n,2] = (idx % x_size, idx / x_size) (coord.py:64-66)
 *   POMDP_COORD_IS_INSIDE : a = coords[n,2]            -> out[n]   = Grid.is_inside (coord.py:61-62, 18-19)
 *   POMDP_COORD_ADD_MOVE  : a = coords[n,2], b = m[n]  -> out[n,2] = Coord + Moves.get_coord(m) (coord.py:10-13, 101-110)
 *   POMDP_COORD_L1        : a, b = coords[n,2]         -> out[n]   = |dx| + |dy| (Grid.euclidean_distance is the
 *                                                                     1-norm, coord.py:79-81)
 *   POMDP_COORD_TAG_GET_INDEX / TAG_GET_COORD / TAG_IS_INSIDE : TagGrid versions (tag.py:46-66);
 *                           TAG_GET_INDEX returns -1 where the reference's asserts would fire.
 * `b` may be NULL for the one-operand ops.                                                 */
#define POMDP_COORD_GET_INDEX      0
#define POMDP_COORD_GET_COORD      1
#define POMDP_COORD_IS_INSIDE      2
#define POMDP_COORD_ADD_MOVE       3
#define POMDP_COORD_L1             4
#define POMDP_COORD_TAG_GET_INDEX  5
#define POMDP_COORD_TAG_GET_COORD  6
#define POMDP_COORD_TAG_IS_INSIDE  7
int pomdp_coord_op(int32_t op, int32_t x_size, int32_t y_size,
                   const int32_t* a, const int32_t* b, int32_t* out, int64_t n, void* stream);

/* ------------------------------------------------------------------ diagnostics ---- */
/* The step kernels' memory behaviour and nothing else (no reference counterpart): per env 4 + 4 bytes read from two
 * streams and 4 x 4 bytes written to four, with the step kernels' grid, vector widths, cache hints and PDL -- no table,
 * no draws, no transition.  bench.py times it next to pomdp_rock_step as the roof of this read:write mix.  All arrays
 * int32/float32 [n], 16-byte aligned, n a multiple of 4; the outputs receive meaningless values.                  */
int pomdp_stream_probe(const int32_t* state, const int32_t* action,
                       int32_t* next_state, int32_t* obs, float* reward, int32_t* flags,
                       int64_t n, void* stream);
/* The same for state_words = 1 or 2 (Rock(15,15): state / next_state int32[2n], moved 256 bits at a time like the step does). */
int pomdp_stream_probe_words(int32_t state_words, const int32_t* state, const int32_t* action,
                             int32_t* next_state, int32_t* obs, float* reward, int32_t* flags,
                             int64_t n, void* stream);

/* ----------------------------------------------------------- belief histogram ------ */
/* Per-shard counts over a batch of packed states, accumulated into int64 hist[bins]
 * (the caller zeroes it and all-reduces it across GPUs with NCCL).  No reference
 * counterpart; bins per env are listed in DESIGN.md:
 *   kind 0 Rock      : [0,k) rocks still good (status +1), then 256 agent cells (x | y<<4)
 *   kind 1 Tag       : 29 agent cells then 29 opponent-0 cells
 *   kind 2 BattleShip: x_size*y_size occupied counts
 *   kind 3 Tiger     : 2 ;  kind 4 Network: n_machines up counts                          */
#define POMDP_KIND_ROCK 0
#define POMDP_KIND_TAG 1
#define POMDP_KIND_BATTLESHIP 2
#define POMDP_KIND_TIGER 3
#define POMDP_KIND_NETWORK 4
int pomdp_belief_hist_bins(int32_t kind, int32_t p0, int32_t p1);
int pomdp_belief_hist(int32_t kind, int32_t p0, int32_t p1, const int32_t* state, int32_t words,
                      int64_t n, long long* hist, void* stream);
/* The same counts in ONE launch, no zero-fill before it: `scratch[POMDP_HIST_MAX_BINS + 2]` int64 is the caller's, zero
 * before the first call.  Every CTA adds its count to a scratch word with ONE atomic that also counts the CTA's arrival
 * at that bin (count in the low 48 bits, arrivals above); the thread that sees the last arrival writes hist_out[bin]
 * (overwritten, not accumulated) and clears the word, so the scratch is all zero again for the next call.  One scratch
 * serves one call at a time (calls on one stream; a second stream needs its own).  n = 0 writes zeros.               */
int pomdp_belief_hist_once(int32_t kind, int32_t p0, int32_t p1, const int32_t* state, int32_t words,
                           int64_t n, long long* scratch, long long* hist_out, void* stream);
/* The same histogram FUSED with its all-reduce over NVLink / NVSwitch peer memory: ONE kernel -- no zero-fill before it,
 * no collective after it.  Every rank passes the same device-resident table `d_peer_bufs[world]` of symmetric buffers,
 * one per rank, each mapped into this process (e.g. torch symmetric memory: _SymmetricMemory.buffer_ptrs_dev).  A
 * buffer is POMDP_HIST_SYMM_WORDS int64: result slot 0 | result slot 1 (POMDP_HIST_MAX_BINS each) | POMDP_HIST_MAX_RANKS
 * arrival counters; all zero before the first call.  `scratch[POMDP_HIST_MAX_BINS + 2]` is this rank's own: zero before
 * the first call, it carries the counts while the CTAs accumulate, a ticket counter and the number of calls made.
 * The CTA that takes the last ticket owns the rank's complete counts; call e = 1, 2, ... uses slot (e - 1) & 1:
 *   1. it clears this rank's other slot (for call e + 1),
 *   2. adds the counts into every rank's slot with system-scope reductions,
 *   3. wait != 0: announces its arrival in every peer's arrival counters (release, system scope) and waits until all
 *      `world` counters of its own buffer have reached e (acquire); a peer that never makes the call trips a 10 s
 *      timeout (the kernel traps) instead of hanging the device.  wait == 0: no signalling (the caller orders a
 *      cross-rank barrier after the kernel and reads the slot itself),
 *   4. copies the slot -- after the wait: the GLOBAL counts -- to hist_out[bins] (may be NULL).
 * The epoch lives in device memory, so the launch arguments never change and the call can be replayed from a CUDA
 * graph.  All ranks must make the same sequence of calls (an empty shard passes n = 0); world <= POMDP_HIST_MAX_RANKS. */
#define POMDP_HIST_MAX_BINS 512
#define POMDP_HIST_MAX_RANKS 64
#define POMDP_HIST_SYMM_WORDS (2 * POMDP_HIST_MAX_BINS + POMDP_HIST_MAX_RANKS)
int pomdp_belief_hist_allreduce(int32_t kind, int32_t p0, int32_t p1, const int32_t* state, int32_t words,
                                int64_t n, long long* scratch, const void* const* d_peer_bufs,
                                int32_t world, int32_t rank, int32_t wait, long long* hist_out, void* stream);

/* The step with the belief histogram of the NEXT states in its epilogue: ONE kernel does
 * pomdp_E_step (same arrays, draws and results) and counts next_state while it is still in registers -- the particle
 * set is not read back from HBM by a second kernel -- and hands the counts to the sink:
 *   d_peer_bufs == NULL : local.  `scratch` as for pomdp_belief_hist_once; hist_out[bins] is overwritten.
 *   d_peer_bufs != NULL : the all-reduce rides in the same launch, exactly as in pomdp_belief_hist_allreduce
 *                         (same scratch, peer table, world, rank, wait, hist_out -- may be NULL -- and call discipline).
 * n = 0 still launches (zeros / the peers are not left waiting).  BattleShip has no such variant (its step kernel moves
 * whole boards with TMA tiles; call pomdp_belief_hist_once after pomdp_battleship_step).                              */
typedef struct {
    long long*         scratch;       /* int64[POMDP_HIST_MAX_BINS + 2], zero before the first call                */
    const void* const* d_peer_bufs;   /* device table of `world` symmetric buffers, or NULL                        */
    int32_t            world, rank, wait;
    int32_t            pad_;
    long long*         hist_out;      /* int64[bins] or NULL (peers only)                                          */
} PomdpHistSink;
int pomdp_rock_step_hist(const PomdpRockParams* params, const void* d_table, const int32_t* state, const int32_t* action,
                         int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n,
                         int64_t global_offset, uint64_t seed, uint32_t step_ctr, const PomdpHistSink* sink, void* stream);
int pomdp_tag_step_hist(const PomdpTagParams* params, const void* d_table, const int32_t* state, const int32_t* action,
                        int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n,
                        int64_t global_offset, uint64_t seed, uint32_t step_ctr, const PomdpHistSink* sink, void* stream);
int pomdp_tiger_step_hist(const PomdpTigerParams* params, const int32_t* state, const int32_t* action,
                          int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n,
                          int64_t global_offset, uint64_t seed, uint32_t step_ctr, const PomdpHistSink* sink, void* stream);
int pomdp_network_step_hist(const PomdpNetworkParams* params, const int32_t* state, const int32_t* action,
                            int32_t* next_state, int32_t* obs, float* reward, int32_t* flags, int64_t n,
                            int64_t global_offset, uint64_t seed, uint32_t step_ctr, const PomdpHistSink* sink, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POMDP_B200_H */
