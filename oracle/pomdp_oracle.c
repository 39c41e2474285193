/*
 * pomdp_oracle.c -- plain-C CPU restatement of d3sm0/gym_pomdp's step()/reset() generative
 * models (RockSample, Tag, BattleShip, Tiger, Network) and Grid/Coord helpers.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load the library built from this file (oracle/_build/
 * libpomdp_oracle.so).  It is the CHECKER for the CUDA kernels, never the product path:
 * nothing under gym_pomdp_b200/ includes, links or dlopens it.
 *
 * Parity status: PINNED.  tests/test_oracle_c_golden.py checks every function here against
 * tests/golden/ (.npz files), which oracle/gen_golden.py recorded from the UNMODIFIED reference
 * (imported from /root/reference through oracle/ref_shim.py, its numpy draws scripted).
 *
 * Deliberately written differently from the kernels (gym_pomdp_b200/csrc/pomdp_core.h):
 * states are UNPACKED arrays in the reference's own units (coordinates, statuses in
 * {-1,0,+1}, per-cell bytes, 0/1 machine flags), probabilities are compared as doubles
 * (u = r / 2^32; binomial(1,p) = u < p) instead of precomputed integer thresholds, boards
 * are byte grids instead of bit masks, and the control flow follows the Python source line
 * by line.  Paths cited below are under /root/reference/gym_pomdp/envs/.
 *
 * Draw coupling (oracle/ref_shim.py): a 32-bit word r stands for one numpy call:
 *   binomial(1,p) -> r/2^32 < p;  uniform() -> r/2^32;  randint(n), choice(len n) -> (r*n)>>32.
 * `draws` arrays are [n_envs, n_slots] uint32, slot tables as in include/pomdp_b200.h; the
 * fill function below produces them with Philox4x32-10 exactly as the kernels do.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TWO32 4294967296.0

/* ------------------------------------------------------------- Philox4x32-10 ------ */
/* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Random123
 * philox4x32 with 10 rounds; constants from the paper. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

void oracle_philox_kat(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    memcpy(out, ctr, 16);
    philox4x32_10(out, key[0], key[1]);
}

/* word(seed, env, step, domain, slot) per include/pomdp_b200.h: one Philox block holds the
 * same slot of four consecutive envs (counter = env >> 2, word = env & 3).  out is [n, n_slots]. */
void oracle_fill_draws(uint64_t seed, uint64_t global_offset, int64_t n, uint32_t step, uint32_t domain,
                       int n_slots, uint32_t* out) {
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t env = global_offset + (uint64_t)i;
        const uint64_t group = env >> 2;
        for (int slot = 0; slot < n_slots; ++slot) {
            uint32_t c[4] = {(uint32_t)group, (uint32_t)(group >> 32), step, (domain << 24) | (uint32_t)slot};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            out[i * n_slots + slot] = c[env & 3];
        }
    }
}

/* Draws keyed by the env itself (BattleShip's fixed-time placement, domain 3): one Philox block holds four consecutive
 * SLOTS of one env -- counter = env, word = slot & 3, block index slot >> 2.  out is [n, n_slots]. */
void oracle_fill_env_draws(uint64_t seed, uint64_t global_offset, int64_t n, uint32_t step, uint32_t domain,
                           int n_slots, uint32_t* out) {
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t env = global_offset + (uint64_t)i;
        for (int slot = 0; slot < n_slots; ++slot) {
            uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), step, (domain << 24) | (uint32_t)(slot >> 2)};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            out[i * n_slots + slot] = c[slot & 3];
        }
    }
}

/* ---- Network: the joint failure draw (include/pomdp_b200.h, "Network draws") ----------------------------------
 * The reference draws binomial(1, p or q) once per machine (network.py:94-99): per machine one uniform against the
 * two thresholds, i.e. three outcomes ("digits": 0 = below both, 1 = between them, 2 = above both).  The kernels
 * sample the 3^5 joint outcomes of five machines from ONE draw word through a 256-column alias table; this is the
 * same table built the same way (integer arithmetic only), and oracle_network_draws() turns the words into
 * per-machine words (0 / lo / hi) that make `word / 2^32 < p` reproduce the digits for oracle_network_step(). */
static uint64_t bern_T(double p) {
    if (!(p > 0.0)) return 0;
    if (p >= 1.0) return 1ull << 32;
    return (uint64_t)ceil(p * TWO32);
}
void oracle_network_alias(double p, double q, uint32_t thr24[256], int32_t alias[256]) {
    const uint64_t Tp = bern_T(p), Tq = bern_T(q), lo = Tp < Tq ? Tp : Tq, hi = Tp < Tq ? Tq : Tp;
    const uint64_t c[3] = {lo, hi - lo, (1ull << 32) - hi};
    uint64_t V[256], S = 0;
    int small[256], large[256], ns = 0, nl = 0;
    for (int k = 0; k < 256; ++k) {
        uint64_t a = 0;
        if (k < 243) {
            a = 1ull << 32;
            for (int i = 0, r = k; i < 5; ++i, r /= 3)
                a = (uint64_t)(((unsigned __int128)a * c[r % 3]) >> 32);
        }
        V[k] = a;
        S += a;
    }
    for (int k = 0; k < 256; ++k) {
        V[k] *= 256;
        alias[k] = k;
        thr24[k] = 0;
        if (V[k] < S) small[ns++] = k; else large[nl++] = k;
    }
    while (ns > 0 && nl > 0) {
        const int s_ = small[--ns], l_ = large[--nl];
        thr24[s_] = (uint32_t)((V[s_] << 24) / S);
        alias[s_] = l_;
        V[l_] -= S - V[s_];
        if (V[l_] < S) small[ns++] = l_; else large[nl++] = l_;
    }
}
/* out: uint32 [N, n_machines + 1] -- per-machine words, then the observation draw's word (slot ceil(n / 5)). */
void oracle_network_draws(uint64_t seed, uint64_t global_offset, int64_t N, uint32_t step, int n_machines, double p,
                          double q, uint32_t* out) {
    uint32_t thr24[256];
    int32_t alias[256];
    const uint64_t Tp = bern_T(p), Tq = bern_T(q), lo = Tp < Tq ? Tp : Tq, hi = Tp < Tq ? Tq : Tp;
    const uint32_t rep[3] = {0u, (uint32_t)lo, hi > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)hi};
    const int G = (n_machines + 4) / 5;
    oracle_network_alias(p, q, thr24, alias);
    for (int64_t i = 0; i < N; ++i) {
        const uint64_t env = global_offset + (uint64_t)i, group = env >> 2;
        uint32_t* o = out + i * (n_machines + 1);
        for (int g = 0; g <= G; ++g) {
            uint32_t cc[4] = {(uint32_t)group, (uint32_t)(group >> 32), step, (uint32_t)g};      /* domain 0 (step) */
            philox4x32_10(cc, (uint32_t)seed, (uint32_t)(seed >> 32));
            const uint32_t w = cc[env & 3];
            if (g == G) { o[n_machines] = w; break; }
            const int col = (int)(w & 255u);
            int code = (alias[col] == col || (uint64_t)w < ((uint64_t)thr24[col] << 8)) ? col : alias[col];
            for (int k = 0; k < 5 && 5 * g + k < n_machines; ++k, code /= 3) o[5 * g + k] = rep[code % 3];
        }
    }
}

static int bern(uint32_t r, double p) { return (double)r / TWO32 < p; }          /* np.random.binomial(1, p) */
static int below(uint32_t r, int n) { return (int)(((uint64_t)r * (uint64_t)n) >> 32); } /* randint(n) */

/* ------------------------------------------------------------------ geometry ------ */
/* coord.py:101-106 Moves: N, E, S, W, NULL */
static const int MOVE_DX[5] = {0, 1, 0, -1, 0};
static const int MOVE_DY[5] = {1, 0, -1, 0, 0};
/* battleship.py:12-21 Compass: N, E, S, W, Null, NE, SE, SW, NW */
static const int COMP_DX[9] = {0, 1, 0, -1, 0, 1, 1, -1, -1};
static const int COMP_DY[9] = {1, 0, -1, 0, 0, 1, -1, -1, 1};

int oracle_grid_get_index(int x_size, int x, int y) { return x_size * y + x; }              /* coord.py:58-59 */
void oracle_grid_get_coord(int x_size, int idx, int* x, int* y) { *x = idx % x_size; *y = idx / x_size; } /* 64-66 */
int oracle_grid_is_inside(int xs, int ys, int x, int y) { return x >= 0 && y >= 0 && x < xs && y < ys; } /* 18-19, 61-62 */
int oracle_l1(int x0, int y0, int x1, int y1) { return abs(x0 - x1) + abs(y0 - y1); }       /* coord.py:79-81 (1-norm) */
void oracle_coord_add_move(int x, int y, int m, int* ox, int* oy) { *ox = x + MOVE_DX[m]; *oy = y + MOVE_DY[m]; }

/* tag.py:46-66 */
int oracle_tag_is_inside(int x, int y) {
    if (y >= 2) return x >= 5 && x < 8 && y < 5;
    return x >= 0 && x < 10 && y >= 0;
}
void oracle_tag_get_coord(int idx, int* x, int* y) {
    if (idx < 20) { *x = idx % 10; *y = idx / 10; return; }
    idx -= 20;
    *x = idx % 3 + 5; *y = idx / 3 + 2;
}
int oracle_tag_get_index(int x, int y) {
    if (y < 2) return y * 10 + x;
    return 20 + (y - 2) * 3 + x - 5;
}

/* ---------------------------------------------------------------- RockSample ------ */
/* rock.py:43-64 */
typedef struct { int board, size_a, size_b, sx, sy, n_listed; int pos[16][2]; } RockConfig;
static const RockConfig ROCK_CONFIGS[] = {
    {2, 2, 1, 0, 0, 1, {{1, 0}}},
    {4, 4, 3, 0, 0, 3, {{1, 0}, {3, 1}, {2, 3}}},
    {7, 7, 8, 0, 3, 8, {{2, 0}, {0, 1}, {3, 1}, {6, 3}, {2, 4}, {3, 4}, {5, 5}, {1, 6}}},
    {11, 11, 11, 0, 5, 11, {{0, 3}, {0, 7}, {1, 8}, {2, 4}, {3, 3}, {3, 8}, {4, 3}, {5, 8}, {6, 1}, {9, 3}, {9, 9}}},
    {15, 15, 15, 0, 5, 16, {{0, 7}, {0, 3}, {1, 2}, {1, 2}, {2, 6}, {3, 7}, {3, 2}, {4, 7}, {5, 2}, {6, 9}, {9, 7},
                            {9, 1}, {11, 8}, {13, 10}, {14, 9}, {12, 2}}},
};
static const RockConfig* rock_config(int board) {
    for (size_t i = 0; i < sizeof(ROCK_CONFIGS) / sizeof(ROCK_CONFIGS[0]); ++i)
        if (ROCK_CONFIGS[i].board == board) return &ROCK_CONFIGS[i];
    return NULL;
}

/* rock.py:106-111: board[x][y] = index of the LAST listed rock at that cell, -1 elsewhere.
 * grid is [n*n] in [x][y] order.  Returns 0, or -1 for an unknown configuration (rock.py:101). */
int oracle_rock_grid(int n, int k, int8_t* grid, int32_t* rock_pos /* [16][2] */, int32_t* start /* [2] */) {
    const RockConfig* c = rock_config(n);
    if (!c || (k != c->size_a && k != c->size_b)) return -1;
    memset(grid, -1, (size_t)n * n);
    for (int i = 0; i < c->n_listed; ++i) {
        grid[c->pos[i][0] * n + c->pos[i][1]] = (int8_t)i;
        rock_pos[2 * i] = c->pos[i][0];
        rock_pos[2 * i + 1] = c->pos[i][1];
    }
    start[0] = c->sx; start[1] = c->sy;
    return c->n_listed;
}

/* rock.py:383-387 */
double oracle_rock_efficiency(int d) { return (1 + pow(2, -(double)d / 20)) * .5; }

/* rock.py:123-194 (RockEnv.step) and rock.py:434-504 (StochasticRockEnv.step), one call per
 * env.  x, y, obs: int32[N]; status: int8[N,k] (updated in place); reward: double[N];
 * done, err: uint8[N]; draws: uint32[N,2] (slot 0 p_move gate, slot 1 sensor).
 * err bit 8 = the reference raises IndexError (grid id >= num_rocks, rock.py:162); defined
 * here as "no rock in this cell". */
int oracle_rock_step(int n, int k, int stochastic, double p_move, int64_t N, int32_t* x, int32_t* y, int8_t* status,
                     const int32_t* action, const uint32_t* draws, int32_t* obs, double* reward, uint8_t* done,
                     uint8_t* err) {
    const RockConfig* c = rock_config(n);
    if (!c) return -1;
    int8_t* grid = (int8_t*)malloc((size_t)n * n);
    int32_t rock_pos[32], start[2];
    if (oracle_rock_grid(n, k, grid, rock_pos, start) < 0) { free(grid); return -1; }
    const int penalization = stochastic ? 0 : -100;               /* rock.py:117, 432 */
    for (int64_t i = 0; i < N; ++i) {
        int8_t* st = status + i * k;
        const int a = action[i];
        int rw = 0, ob = 0, fin = 0;
        err[i] = 0;
        if (stochastic && !bern(draws[2 * i], p_move)) {            /* rock.py:443 */
            obs[i] = 0; reward[i] = 0; done[i] = 0;
            continue;
        }
        if (a < 4) {
            if (a == 1) {                                           /* EAST, rock.py:135-141 */
                if (x[i] + 1 < n) x[i] += 1;
                else { rw = 10; fin = 1; }
            } else if (a == 0) {                                    /* NORTH */
                if (y[i] + 1 < n) y[i] += 1; else rw = penalization;
            } else if (a == 2) {                                    /* SOUTH */
                if (y[i] - 1 >= 0) y[i] -= 1; else rw = penalization;
            } else {                                                /* WEST */
                if (x[i] - 1 >= 0) x[i] -= 1; else rw = penalization;
            }
        } else if (a == 4) {                                        /* SAMPLE, rock.py:160-169 */
            int rock = grid[x[i] * n + y[i]];
            if (rock >= k) { err[i] |= 8; rock = -1; }
            if (rock >= 0 && st[rock] != 0) {
                rw = st[rock] == 1 ? 10 : -10;
                st[rock] = 0;
            } else {
                rw = penalization;
            }
        } else {                                                    /* CHECK, rock.py:171-175, 401-407 */
            const int rock = a - 4 - 1;
            const double eff = oracle_rock_efficiency(oracle_l1(x[i], y[i], rock_pos[2 * rock], rock_pos[2 * rock + 1]));
            if (bern(draws[2 * i + 1], eff)) ob = st[rock] == 1 ? 2 : 1;
            else ob = st[rock] == 1 ? 1 : 2;
        }
        if (!fin && !stochastic) fin = (penalization == rw);        /* rock.py:193; commented out at 503 */
        obs[i] = ob; reward[i] = rw; done[i] = (uint8_t)fin;
    }
    free(grid);
    return 0;
}

/* rock.py:236-241, 266-271, 78-80: status = int(np.sign(uniform(0,1) - .5)), one uniform per rock.  Rock i's
 * uniform is r_i / 2^32 with r_i = rotl32(word of reset slot 0, 30 - 2 i); draws is [N, 1]. */
int oracle_rock_reset(int n, int k, int64_t N, const uint32_t* draws, int32_t* x, int32_t* y,
                      int8_t* status, int32_t* obs) {
    const RockConfig* c = rock_config(n);
    if (!c) return -1;
    for (int64_t i = 0; i < N; ++i) {
        x[i] = c->sx; y[i] = c->sy; obs[i] = 0;
        for (int r = 0; r < k; ++r) {
            const uint32_t w = draws[i];
            const int sh = (30 - 2 * r) & 31;
            const uint32_t ri = sh ? (w << sh) | (w >> (32 - sh)) : w;
            const double v = (double)ri / TWO32 - .5;
            status[i * k + r] = (int8_t)((v > 0) - (v < 0));
        }
    }
    return 0;
}

/* --------------------------------------------------------------------- Tag ------ */
/* tag.py:260-280 */
static int tag_admissible(int ax, int ay, int ox, int oy, int acts[8]) {
    int n = 0;
    if (ox >= ax) acts[n++] = 1;
    if (oy >= ay) acts[n++] = 0;
    if (ox <= ax) acts[n++] = 3;
    if (oy <= ay) acts[n++] = 2;
    if (ox == ax && oy > ay) acts[n++] = 0;
    if (oy == ay && ox > ax) acts[n++] = 1;
    if (ox == ax && oy < ay) acts[n++] = 2;
    if (oy == ay && ox < ax) acts[n++] = 3;
    return n;
}
void oracle_tag_admissible(int agent, int opp, int8_t out[4]) {
    int ax, ay, ox, oy, acts[8];
    oracle_tag_get_coord(agent, &ax, &ay);
    oracle_tag_get_coord(opp, &ox, &oy);
    const int n = tag_admissible(ax, ay, ox, oy, acts);
    for (int j = 0; j < 4; ++j) out[j] = j < n ? (int8_t)acts[j] : -1;
}

/* tag.py:108-143 (+ move_opponent 201-207, _sample_ob 219-226).  agent: int32[N] cell ids,
 * opp: int32[N,n_opp], num_opp: int32[N] (all updated in place); draws uint32[N, n_opp]: slot j is opponent j's
 * word -- binomial(1, move_prob) reads it whole, choice(actions) its low half (the word w << 16). */
void oracle_tag_step(int n_opp, double move_prob, int64_t N, int32_t* agent, int32_t* opp, int32_t* num_opp,
                     const int32_t* action, const uint32_t* draws, int32_t* obs, double* reward, uint8_t* done) {
    for (int64_t i = 0; i < N; ++i) {
        int ax, ay;
        oracle_tag_get_coord(agent[i], &ax, &ay);
        int32_t* o = opp + i * n_opp;
        const uint32_t* dr = draws + i * n_opp;
        const int a = action[i];
        double rw = 0.;
        if (a == 4) {
            int tagged = 0;
            for (int j = 0; j < n_opp; ++j) {
                int ox, oy;
                oracle_tag_get_coord(o[j], &ox, &oy);
                if (ox == ax && oy == ay) {
                    rw = 10.; tagged = 1; num_opp[i] -= 1;
                } else if (oracle_tag_is_inside(ox, oy) && num_opp[i] > 0) {
                    int acts[8];
                    const int cnt = tag_admissible(ax, ay, ox, oy, acts);      /* tag.py:203 */
                    if (bern(dr[j], move_prob)) {                              /* tag.py:204 */
                        const int m = acts[below(dr[j] << 16, cnt)];           /* tag.py:205 */
                        if (oracle_tag_is_inside(ox + MOVE_DX[m], oy + MOVE_DY[m]))
                            o[j] = oracle_tag_get_index(ox + MOVE_DX[m], oy + MOVE_DY[m]);
                    }
                }
            }
            if (!tagged) rw = -10.;
        } else {
            rw = -1.;
            if (oracle_tag_is_inside(ax + MOVE_DX[a], ay + MOVE_DY[a])) { ax += MOVE_DX[a]; ay += MOVE_DY[a]; }
            agent[i] = oracle_tag_get_index(ax, ay);
        }
        int ob = agent[i];                                                     /* tag.py:219-226 */
        if (a < 4)
            for (int j = 0; j < n_opp; ++j)
                if (o[j] == agent[i]) ob = 29;
        obs[i] = ob; reward[i] = rw; done[i] = num_opp[i] == 0;                 /* tag.py:142 */
    }
}

/* tag.py:97-102, 181-193, 43-44: 1 + n_opp calls of randint(29), agent first; ob = _sample_ob(state, 0).  The j-th
 * call is scripted with the word (slot j / 3) * 29^(j % 3) mod 2^32, i.e. it returns the (j % 3)-th base-29 digit of
 * that slot's uniform; draws is [N, (1 + n_opp + 2) / 3]. */
static uint32_t tag_reset_word(uint32_t w, int digit) { return digit == 0 ? w : digit == 1 ? w * 29u : w * 841u; }
void oracle_tag_reset(int n_opp, int64_t N, const uint32_t* draws, int32_t* agent, int32_t* opp,
                      int32_t* num_opp, int32_t* obs) {
    const int n_slots = (1 + n_opp + 2) / 3;
    for (int64_t i = 0; i < N; ++i) {
        const uint32_t* dr = draws + i * n_slots;
        agent[i] = below(tag_reset_word(dr[0], 0), 29);
        obs[i] = agent[i];
        for (int j = 0; j < n_opp; ++j) {
            opp[i * n_opp + j] = below(tag_reset_word(dr[(1 + j) / 3], (1 + j) % 3), 29);
            if (opp[i * n_opp + j] == agent[i]) obs[i] = 29;
        }
        num_opp[i] = n_opp;
    }
}

/* -------------------------------------------------------------- BattleShip ------ */
/* Boards are uint8 [xs*ys] in the reference's board[x][y] order (battleship.py:57-61). */
static int ship_collision(int xs, int ys, const uint8_t* occ, int x, int y, int dir, int length) {
    for (int i = 0; i < length + 1; ++i) {                                      /* battleship.py:198 */
        if (!oracle_grid_is_inside(xs, ys, x + COMP_DX[dir], y + COMP_DY[dir])) return 1;
        if (occ[x * ys + y]) return 1;
        for (int adj = 0; adj < 8; ++adj) {                                     /* range(8): NW never looked at */
            const int cx = x + COMP_DX[adj], cy = y + COMP_DY[adj];
            if (oracle_grid_is_inside(xs, ys, cx, cy) && occ[cx * ys + cy]) return 1;
        }
        x += COMP_DX[dir]; y += COMP_DY[dir];
    }
    return 0;
}
static void ship_mark(int ys, uint8_t* occ, int x, int y, int dir, int length) {   /* battleship.py:182-193 */
    for (int i = 0; i < length; ++i) {
        occ[x * ys + y] = 1;
        x += COMP_DX[dir]; y += COMP_DY[dir];
    }
}

/* battleship.py:131-137, 167-180 as written: rejection loop; attempt a uses slot 2a =
 * randint(n_tiles) (coord.py:68) and 2a+1 = randint(4) (battleship.py:36).  occ: uint8[N,xs*ys];
 * ships: int32[N, n_ships, 4] = (x, y, dir, length); attempts: int32[N] (-1: ran out of slots). */
void oracle_battleship_reset_rejection(int xs, int ys, int max_len, int64_t N, const uint32_t* draws, int n_slots,
                                       uint8_t* occ, int32_t* ships, int32_t* attempts, int32_t* remaining) {
    const int n_tiles = xs * ys, n_ships = max_len - 1;
    for (int64_t i = 0; i < N; ++i) {
        uint8_t* o = occ + i * n_tiles;
        memset(o, 0, (size_t)n_tiles);
        int a = 0, s = 0, total = 0, fail = 0;
        for (int length = max_len; length >= 2 && !fail; --length, ++s) {        /* reversed(range(2, max_len+1)) */
            for (;;) {
                if (2 * a + 1 >= n_slots) { fail = 1; break; }
                const int pos = below(draws[i * n_slots + 2 * a], n_tiles);
                const int dir = below(draws[i * n_slots + 2 * a + 1], 4);
                ++a;
                const int x = pos % xs, y = pos / xs;                            /* coord.py:64-66 */
                if (!ship_collision(xs, ys, o, x, y, dir, length)) {
                    ship_mark(ys, o, x, y, dir, length);
                    int32_t* sh = ships + (i * n_ships + s) * 4;
                    sh[0] = x; sh[1] = y; sh[2] = dir; sh[3] = length;
                    total += length;
                    break;
                }
            }
        }
        attempts[i] = fail ? -1 : a;
        remaining[i] = total;
    }
}

/* The set the rejection loop samples uniformly from: valid[c], c = 4*pos + dir. Returns its size. */
int oracle_battleship_valid(int xs, int ys, const uint8_t* occ, int length, uint8_t* valid) {
    int cnt = 0;
    for (int c = 0; c < 4 * xs * ys; ++c) {
        const int pos = c >> 2, dir = c & 3;
        valid[c] = !ship_collision(xs, ys, occ, pos % xs, pos / xs, dir, length);
        cnt += valid[c];
    }
    return cnt;
}

/* The fixed-time equivalent the warp kernel uses: ship s = the floor(u*count)-th valid candidate
 * in increasing c, u from slot s. */
void oracle_battleship_reset_scan(int xs, int ys, int max_len, int64_t N, const uint32_t* draws, int n_slots,
                                  uint8_t* occ, int32_t* remaining, uint8_t* err) {
    const int n_tiles = xs * ys;
    uint8_t* valid = (uint8_t*)malloc((size_t)4 * n_tiles);
    for (int64_t i = 0; i < N; ++i) {
        uint8_t* o = occ + i * n_tiles;
        memset(o, 0, (size_t)n_tiles);
        int s = 0, total = 0;
        err[i] = 0;
        for (int length = max_len; length >= 2; --length, ++s) {
            const int cnt = oracle_battleship_valid(xs, ys, o, length, valid);
            if (cnt == 0) { err[i] = 8; break; }
            int kth = below(draws[i * n_slots + s], cnt);
            int c = 0;
            for (;; ++c) if (valid[c] && kth-- == 0) break;
            ship_mark(ys, o, (c >> 2) % xs, (c >> 2) / xs, c & 3, length);
            total += length;
        }
        remaining[i] = total;
    }
    free(valid);
}

/* battleship.py:91-122.  occ, vis: uint8[N, xs*ys] ([x][y] order); remaining int32[N]. */
void oracle_battleship_step(int xs, int ys, int64_t N, const uint8_t* occ, uint8_t* vis, int32_t* remaining,
                            const int32_t* action, int32_t* obs, double* reward, uint8_t* done) {
    const int n_tiles = xs * ys;
    for (int64_t i = 0; i < N; ++i) {
        const int x = action[i] % xs, y = action[i] / xs;                       /* coord.py:64-66 */
        const int cell = x * ys + y;
        int rw = 0, ob = 0;
        if (vis[i * n_tiles + cell]) {
            rw -= 10;
        } else {
            if (occ[i * n_tiles + cell]) { rw -= 1; ob = 1; remaining[i] -= 1; }
            else rw -= 1;
            vis[i * n_tiles + cell] = 1;
        }
        done[i] = 0;
        if (remaining[i] == 0) { rw += n_tiles; done[i] = 1; }                  /* battleship.py:118-120 */
        obs[i] = ob; reward[i] = rw;
    }
}

/* ------------------------------------------------------------------- Tiger ------ */
/* tiger.py:72-88, 117-119, 140-172.  draws: uint32[N], ONE word per env: state_space.sample() after an OPEN and uniform()
 * after a LISTEN read the same word (a step never consumes both). */
void oracle_tiger_step(double listen_prob, int64_t N, int32_t* state, const int32_t* action, const uint32_t* draws,
                       int32_t* obs, double* reward, uint8_t* done) {
    for (int64_t i = 0; i < N; ++i) {
        const int a = action[i];
        const int terminal = a != 2 && ((a == 0 && state[i] == 0) || (a == 1 && state[i] == 1));   /* 155-162 */
        reward[i] = a == 2 ? -1 : (!terminal ? 10 : -20);                                          /* 164-172 */
        if (terminal) { done[i] = 1; obs[i] = state[i]; continue; }                               /* 81-83 */
        if (a == 1 || a == 0) state[i] = below(draws[i], 2);                                       /* 117-119 */
        const double p = (double)draws[i] / TWO32;                 /* 143: the same word; read only when a == 2 */
        int ob = 2;
        if (a == 2) {
            if (state[i] == 0) ob = p > listen_prob ? 1 : 0;
            else ob = p > listen_prob ? 0 : 1;
        }
        obs[i] = ob; done[i] = 0;
    }
}
void oracle_tiger_reset(int64_t N, const uint32_t* draws, int32_t* state, int32_t* obs) {   /* tiger.py:60-66 */
    for (int64_t i = 0; i < N; ++i) { state[i] = below(draws[i], 2); obs[i] = 2; }
}

/* ----------------------------------------------------------------- Network ------ */
/* network.py:144-168.  nb: int32[n,3] padded with -1; returns 0 or -1 (assert at 155). */
int oracle_network_neighbours(int n, int problem_type, int32_t* nb) {
    int cnt[64] = {0};
    for (int i = 0; i < 3 * n; ++i) nb[i] = -1;
#define LINK(i, j) nb[3 * (i) + cnt[i]++] = (j)
    if (problem_type == 3) {
        if (n < 4 || n % 3 != 1) return -1;
        LINK(0, 1); LINK(0, 2); LINK(0, 3);
        for (int i = 1; i < n; ++i) {
            if (i < n - 3) LINK(i, i + 3);
            if (i <= 4) LINK(i, 0); else LINK(i, i - 3);
        }
    } else {
        for (int i = 0; i < n; ++i) { LINK(i, (i + 1) % n); LINK(i, (i + n - 1) % n); }
    }
#undef LINK
    return 0;
}

/* network.py:71-114.  machines: int8[N,n] (in place); draws uint32[N, n+1]; reward double. */
int oracle_network_step(int n, int problem_type, double p, double q, double p_ob, int64_t N, int8_t* machines,
                        const int32_t* action, const uint32_t* draws, int32_t* obs, double* reward) {
    int32_t nb[3 * 64];
    if (n > 64 || oracle_network_neighbours(n, problem_type, nb)) return -1;
    for (int64_t i = 0; i < N; ++i) {
        int8_t* s = machines + i * n;
        const uint32_t* dr = draws + i * (n + 1);
        int8_t n_fail[64] = {0};
        double rw = 0;
        int ob = 2;
        for (int m = 0; m < n; ++m)
            for (int j = 0; j < 3 && nb[3 * m + j] >= 0; ++j)
                if (s[nb[3 * m + j]] == 0) n_fail[m] = 1;
        for (int m = 0; m < n; ++m)
            if (s[m] == 1) rw += (nb[3 * m + 2] >= 0) ? 2 : 1;                  /* len(neighbours) > 2 */
        for (int m = 0; m < n; ++m)
            if (s[m]) s[m] = (int8_t)(1 - bern(dr[m], n_fail[m] ? q : p));
        const int a = action[i];
        if (a < 2 * n) {
            const int machine = a / 2, reboot = a % 2;
            if (reboot) {
                rw -= 2.5;
                s[machine] = 1;
                ob = bern(dr[n], p_ob);
            } else {
                rw -= .1;
                ob = bern(dr[n], p_ob) ? s[machine] : 1 - s[machine];
            }
        }
        obs[i] = ob; reward[i] = rw;
    }
    return 0;
}

/* ------------------------------------- legal actions + uniform-legal rollouts ------ */
/* SURVEY.md §8f rank 1.  The loops at rock.py:563-572 and tag.py:310-316:
 *     a = np.random.choice(env._generate_legal()); ob, rw, done, _ = env.step(a);
 *     r += rw * discount; discount *= env._discount
 * Rollout step t of env e uses the Philox words of step counter ctr0 + t: domain 2 (POLICY)
 * slot 0 for `choice` (index = floor(u * len)), domain 0 (STEP) for the step's own slots. */
static uint32_t draw_word1(uint64_t seed, uint64_t env, uint32_t step, uint32_t domain, int slot) {
    const uint64_t group = env >> 2;
    uint32_t c[4] = {(uint32_t)group, (uint32_t)(group >> 32), step, (domain << 24) | (uint32_t)slot};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return c[env & 3];
}

/* rock.py:273-291, list order preserved.  The dangling cell (grid id >= k), where the
 * reference's _generate_legal itself raises IndexError, offers no SAMPLE. */
int oracle_rock_legal(int n, int k, int x, int y, const int8_t* status, int32_t* legal) {
    int8_t* grid = (int8_t*)malloc((size_t)n * n);
    int32_t rock_pos[32], start[2];
    int cnt = 0;
    if (oracle_rock_grid(n, k, grid, rock_pos, start) < 0) { free(grid); return -1; }
    legal[cnt++] = 1;
    if (y + 1 < n) legal[cnt++] = 0;
    if (y - 1 >= 0) legal[cnt++] = 2;
    if (x - 1 >= 0) legal[cnt++] = 3;
    {
        const int rock = grid[x * n + y];
        if (rock >= 0 && rock < k && status[rock] != 0) legal[cnt++] = 4;
    }
    for (int r = 0; r < k; ++r)
        if (status[r] != 0) legal[cnt++] = grid[rock_pos[2 * r] * n + rock_pos[2 * r + 1]] + 1 + 4;
    free(grid);
    return cnt;
}

int oracle_rock_rollout(int n, int k, int stochastic, double p_move, int64_t N, int32_t* x, int32_t* y, int8_t* status,
                        uint64_t seed, uint64_t goff, uint32_t ctr0, int max_steps, double gamma, const int32_t* first_action, double* ret,
                        int32_t* steps, uint8_t* done, uint8_t* err) {
    for (int64_t i = 0; i < N; ++i) {
        double r = 0., disc = 1.;
        int t = 0;
        uint8_t fin = 0, e_acc = 0;
        while (t < max_steps && !fin) {
            int32_t legal[40], a, ob;
            uint32_t dr[2];
            double rw;
            uint8_t e1;
            const int cnt = oracle_rock_legal(n, k, x[i], y[i], status + i * k, legal);
            if (cnt < 0) return -1;
            a = legal[below(draw_word1(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, 2, 0), cnt)];
            if (t == 0 && first_action) a = first_action[i];           /* the caller's action first: Q(s, a) */
            dr[0] = draw_word1(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, 0, 0);
            dr[1] = draw_word1(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, 0, 1);
            if (oracle_rock_step(n, k, stochastic, p_move, 1, x + i, y + i, status + i * k, &a, dr, &ob, &rw, &fin, &e1))
                return -1;
            r += rw * disc;
            disc *= gamma;
            e_acc |= e1;
            ++t;
        }
        ret[i] = r; steps[i] = t; done[i] = fin; err[i] = e_acc;
    }
    return 0;
}

void oracle_tag_rollout(int n_opp, double move_prob, int64_t N, int32_t* agent, int32_t* opp, int32_t* num_opp,
                        uint64_t seed, uint64_t goff, uint32_t ctr0, int max_steps, double gamma, const int32_t* first_action, double* ret,
                        int32_t* steps, uint8_t* done) {
    for (int64_t i = 0; i < N; ++i) {
        double r = 0., disc = 1.;
        int t = 0;
        uint8_t fin = 0;
        while (t < max_steps && !fin) {
            uint32_t dr[8];
            int32_t a = below(draw_word1(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, 2, 0), 5), ob;   /* tag.py:228-229 */
            double rw;
            if (t == 0 && first_action) a = first_action[i];
            for (int s = 0; s < n_opp; ++s) dr[s] = draw_word1(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, 0, s);
            oracle_tag_step(n_opp, move_prob, 1, agent + i, opp + i * n_opp, num_opp + i, &a, dr, &ob, &rw, &fin);
            r += rw * disc;
            disc *= gamma;
            ++t;
        }
        ret[i] = r; steps[i] = t; done[i] = fin;
    }
}

void oracle_tiger_rollout(double listen_prob, int64_t N, int32_t* state, uint64_t seed, uint64_t goff, uint32_t ctr0,
                          int max_steps, double gamma, const int32_t* first_action, double* ret, int32_t* steps, uint8_t* done) {
    for (int64_t i = 0; i < N; ++i) {
        double r = 0., disc = 1.;
        int t = 0;
        uint8_t fin = 0;
        while (t < max_steps && !fin) {
            uint32_t dr[1];
            int32_t a = below(draw_word1(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, 2, 0), 3), ob;   /* tiger.py:111-112 */
            double rw;
            if (t == 0 && first_action) a = first_action[i];
            dr[0] = draw_word1(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, 0, 0);
            oracle_tiger_step(listen_prob, 1, state + i, &a, dr, &ob, &rw, &fin);
            r += rw * disc;
            disc *= gamma;
            ++t;
        }
        ret[i] = r; steps[i] = t; done[i] = fin;
    }
}

int oracle_network_rollout(int n, int problem_type, double p, double q, double p_ob, int64_t N, int8_t* machines,
                           uint64_t seed, uint64_t goff, uint32_t ctr0, int max_steps, double gamma, const int32_t* first_action, double* ret,
                           int32_t* steps) {
    for (int64_t i = 0; i < N; ++i) {
        double r = 0., disc = 1.;
        for (int t = 0; t < max_steps; ++t) {                                                            /* never done */
            uint32_t dr[65];
            int32_t a = below(draw_word1(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, 2, 0), 2 * n + 1), ob;   /* network.py:129-130 */
            double rw;
            if (t == 0 && first_action) a = first_action[i];
            oracle_network_draws(seed, goff + (uint64_t)i, 1, ctr0 + (uint32_t)t, n, p, q, dr);
            if (oracle_network_step(n, problem_type, p, q, p_ob, 1, machines + i * n, &a, dr, &ob, &rw)) return -1;
            r += rw * disc;
            disc *= gamma;
        }
        ret[i] = r; steps[i] = max_steps;
    }
    return 0;
}

/* battleship.py:157-165: legal = unvisited cells in increasing action order */
void oracle_battleship_rollout(int xs, int ys, int64_t N, const uint8_t* occ, uint8_t* vis, int32_t* remaining,
                               uint64_t seed, uint64_t goff, uint32_t ctr0, int max_steps, double gamma, const int32_t* first_action, double* ret,
                               int32_t* steps, uint8_t* done) {
    const int n_tiles = xs * ys;
    for (int64_t i = 0; i < N; ++i) {
        double r = 0., disc = 1.;
        int t = 0;
        uint8_t fin = 0;
        while (t < max_steps && !fin) {
            int32_t legal[128], a, ob;
            int cnt = 0;
            double rw;
            for (int act = 0; act < n_tiles; ++act) {
                const int x = act % xs, y = act / xs;
                if (!vis[i * n_tiles + x * ys + y]) legal[cnt++] = act;
            }
            a = cnt ? legal[below(draw_word1(seed, goff + (uint64_t)i, ctr0 + (uint32_t)t, 2, 0), cnt)] : 0;
            if (t == 0 && first_action) a = first_action[i];
            oracle_battleship_step(xs, ys, 1, occ + i * n_tiles, vis + i * n_tiles, remaining + i, &a, &ob, &rw, &fin);
            r += rw * disc;
            disc *= gamma;
            ++t;
        }
        ret[i] = r; steps[i] = t; done[i] = fin;
    }
}
